// tcgen05 / TMEM / TMA implicit-GEMM conv with the Conv1dBlock + residual-block epilogue fused (K2 in SURVEY.md
// Appendix C).  Replaces the same reference statements as conv_ffma.cu (modeling/helpers.py:95-112,
// modeling/temporal.py:53-55, 227, 233-245) for every layer whose channel counts are multiples of 64.
//
//   D[rows, Cout] = sum_{tap, c} A_tap[rows, c] * W_tap[Cout, c]^T        rows = (sample, position), channels-last
//
// * A (activations) and B (weights) are bf16, K-major, staged by TMA into 128B-swizzled shared memory.  The conv taps
//   are row-shifted TMA boxes of the SAME 3-D tensor map (C, L, B): the L coordinate starts at (tap - pad) and the
//   hardware zero-fills out-of-range positions, so padding costs nothing and there is no im2col.
// * precision modes: NSPLIT = 1 -> single bf16 pass;  NSPLIT = 2 -> activations and weights are stored as bf16 hi/lo
//   pairs and three MMAs (hi*hi + lo*hi + hi*lo) accumulate in fp32 in TMEM ("bf16x3", fp32-class parity).
// * one elected thread issues tcgen05.mma (cta_group::1, M=128, N=64, K=16); accumulators live in TMEM
//   (columns [0,64) main GEMM, [64,128) the residual 1x1 conv of the block input when present).
// * epilogue: thread t owns TMEM lane t == tile row t: bias, GroupNorm(8) statistics by warp shuffles over the L rows
//   of a sample, Mish, + time embedding, + residual, optional fused 1x1 head, bf16 hi/lo store.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue (TMEM lane quadrants
// 2,3,0,1).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace b2p {

constexpr int TC_M = 128;           // rows per tile
constexpr int TC_N = 64;            // output channels per tile
constexpr int TC_K = 64;            // channels per pipeline stage (128 bytes of bf16: one swizzle atom row)
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 128;
constexpr int A_BYTES = TC_M * TC_K * 2;   // 16 KB
constexpr int B_BYTES = TC_N * TC_K * 2;   //  8 KB

template <int NSPLIT> struct StageBytes { static constexpr int value = NSPLIT * (A_BYTES + B_BYTES); };
template <int NSPLIT> struct NumStages { static constexpr int value = NSPLIT == 2 ? 4 : 6; };

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B swizzle, 8-row groups 1024 B apart (sm100 descriptor version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
  d |= (uint64_t)1 << 46;                            // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
// kind::f16: D fp32, A/B bf16, both K-major, M=128, N=64
__device__ __forceinline__ constexpr uint32_t umma_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// GroupNorm(8) + Mish on one tile row held in registers: a group is CG consecutive channels x the L rows (lanes) of a
// sample; per-thread partial sums are combined across the L lanes with xor shuffles (two-pass: mean, then variance).
template <int CG>
__device__ __forceinline__ void group_norm_mish(float (&v)[TC_N], int L, const float* __restrict__ gamma, const float* __restrict__ beta) {
  const float inv_n = 1.0f / (float)(CG * L);
#pragma unroll
  for (int g0 = 0; g0 < TC_N; g0 += CG) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CG; ++c) s += v[g0 + c];
    for (int o = 1; o < L; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_n;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CG; ++c) { float d = v[g0 + c] - mean; q = fmaf(d, d, q); }
    for (int o = 1; o < L; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * inv_n + 1e-5f);
#pragma unroll
    for (int c = 0; c < CG; ++c) v[g0 + c] = mish_f((v[g0 + c] - mean) * rstd * __ldg(gamma + g0 + c) + __ldg(beta + g0 + c));
  }
}

struct __align__(16) TcBarriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full;
  uint32_t tmem_base;
  uint32_t pad;
};

template <int NSPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcArgs a) {
  constexpr int STAGES = NumStages<NSPLIT>::value;
  constexpr int STAGE_BYTES = StageBytes<NSPLIT>::value;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem + STAGES * STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, n0 = blockIdx.y * TC_N, par = blockIdx.z;
  const int b0 = tile_m * a.samples_per_tile;

  // iteration space of the K loop: main phase (taps x sources x 64-channel chunks), then the residual 1x1 phase
  const int chunks0 = a.C[0] / TC_K, chunks1 = a.C[1] / TC_K;
  const int main_iters = a.ntaps * (chunks0 + chunks1);
  const int rchunks0 = a.RC[0] / TC_K, rchunks1 = a.RC[1] / TC_K;
  const int res_iters = rchunks0 + rchunks1;
  const int total_iters = main_iters + res_iters;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    mbar_init(&bars->tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation (one warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      prefetch_tmap(&maps.a[0][0]);
      prefetch_tmap(&maps.w[0]);
      for (int it = 0; it < total_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&bars->empty[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&bars->full[s], STAGE_BYTES);
        if (it < main_iters) {
          const int per_tap = chunks0 + chunks1;
          const int tap = it / per_tap, ch = it % per_tap;
          const int src = ch < chunks0 ? 0 : 1;
          const int c0 = (src == 0 ? ch : ch - chunks0) * TC_K;
          const int kglob = (src == 0 ? 0 : a.C[0]) + c0;                     // column in the packed weight matrix
          const int l0 = a.tap_l0[par][tap];
          const int wrow = a.tap_w[par][tap] * a.Cout + n0;
#pragma unroll
          for (int h = 0; h < NSPLIT; ++h) {
            tma_load_3d(st + h * A_BYTES, &maps.a[src][h], &bars->full[s], c0, l0, b0);
            tma_load_2d(st + NSPLIT * A_BYTES + h * B_BYTES, &maps.w[h], &bars->full[s], kglob, wrow);
          }
        } else {
          const int ch = it - main_iters;
          const int src = ch < rchunks0 ? 0 : 1;
          const int c0 = (src == 0 ? ch : ch - rchunks0) * TC_K;
          const int kglob = (src == 0 ? 0 : a.RC[0]) + c0;
#pragma unroll
          for (int h = 0; h < NSPLIT; ++h) {
            tma_load_3d(st + h * A_BYTES, &maps.r[src][h], &bars->full[s], c0, 0, b0);
            tma_load_2d(st + NSPLIT * A_BYTES + h * B_BYTES, &maps.rw[h], &bars->full[s], kglob, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc();
      for (int it = 0; it < total_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&bars->full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + NSPLIT * A_BYTES;
        const bool res_phase = it >= main_iters;
        const uint32_t d = tmem_base + (res_phase ? TC_N : 0);
        const bool first = res_phase ? (it == main_iters) : (it == 0);
#pragma unroll
        for (int k = 0; k < TC_K / TC_UMMA_K; ++k) {
          const uint32_t koff = k * TC_UMMA_K * 2;   // bytes inside the 128B swizzle row
          const uint64_t a_hi = umma_desc(sa + koff), b_hi = umma_desc(sb + koff);
          umma(d, a_hi, b_hi, idesc, (first && k == 0) ? 0u : 1u);
          if (NSPLIT == 2) {
            const uint64_t a_lo = umma_desc(sa + A_BYTES + koff), b_lo = umma_desc(sb + B_BYTES + koff);
            umma(d, a_lo, b_hi, idesc, 1u);
            umma(d, a_hi, b_lo, idesc, 1u);
          }
        }
        umma_commit(&bars->empty[s]);          // frees the smem stage when these MMAs retire
      }
      umma_commit(&bars->tmem_full);           // accumulators complete
    }
  } else {
    // =============================== epilogue (warps 2..5) ===============================
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;                 // tile row == TMEM lane
    const int L = a.Lrows;
    const long grow = (long)tile_m * TC_M + r;
    const bool ok = grow < a.nrows;
    const int b = (int)(grow >> a.log2L), l = (int)(grow & (L - 1));
    mbar_wait(&bars->tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float v[TC_N];
    tmem_ld32(taddr, v);
    tmem_ld32(taddr + 32, v + 32);
#pragma unroll
    for (int c = 0; c < TC_N; ++c) v[c] += __ldg(a.bias + n0 + c);
    if (a.gn_gamma) {
      switch (a.cg) {    // compile-time group width keeps v[] in registers
        case 8: group_norm_mish<8>(v, L, a.gn_gamma + n0, a.gn_beta + n0); break;
        case 16: group_norm_mish<16>(v, L, a.gn_gamma + n0, a.gn_beta + n0); break;
        case 32: group_norm_mish<32>(v, L, a.gn_gamma + n0, a.gn_beta + n0); break;
        default: group_norm_mish<64>(v, L, a.gn_gamma + n0, a.gn_beta + n0); break;
      }
    }
    if (a.temb && ok) {
      const float* t = a.temb + (size_t)b * a.temb_stride + n0;
#pragma unroll
      for (int c = 0; c < TC_N; c += 4) { float4 t4 = __ldg(reinterpret_cast<const float4*>(t + c)); v[c] += t4.x; v[c + 1] += t4.y; v[c + 2] += t4.z; v[c + 3] += t4.w; }
    }
    if (res_iters > 0) {                             // residual 1x1 conv accumulated in TMEM columns [64,128)
      float rv[32];
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        tmem_ld32(taddr + TC_N + 32 * hlf, rv);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[32 * hlf + c] += rv[c] + __ldg(a.resB + n0 + 32 * hlf + c);
      }
    }
    if (ok) {
      if (a.res_f32) {
        const float* q = a.res_f32 + (size_t)grow * a.Cout + n0;
#pragma unroll
        for (int c = 0; c < TC_N; c += 4) { float4 t4 = __ldg(reinterpret_cast<const float4*>(q + c)); v[c] += t4.x; v[c + 1] += t4.y; v[c + 2] += t4.z; v[c + 3] += t4.w; }
      }
      if (a.res_hi) {
        const uint4* qh = reinterpret_cast<const uint4*>(a.res_hi + (size_t)grow * a.Cout + n0);
        const uint4* ql = a.res_lo ? reinterpret_cast<const uint4*>(a.res_lo + (size_t)grow * a.Cout + n0) : nullptr;
#pragma unroll
        for (int c8 = 0; c8 < TC_N / 8; ++c8) {
          uint4 hh = __ldg(qh + c8);
          const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&hh);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[c8 * 8 + i] += __bfloat162float(hp[i]);
          if (ql) {
            uint4 ll = __ldg(ql + c8);
            const __nv_bfloat16* lp = reinterpret_cast<const __nv_bfloat16*>(&ll);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[c8 * 8 + i] += __bfloat162float(lp[i]);
          }
        }
      }
      const size_t orow = (size_t)b * a.out_L + (size_t)l * a.out_lstride + (par ? a.out_loff1 : a.out_loff0);
      if (a.out_hi) {
        uint4* oh = reinterpret_cast<uint4*>(a.out_hi + orow * a.Cout + n0);
        uint4* ol = a.out_lo ? reinterpret_cast<uint4*>(a.out_lo + orow * a.Cout + n0) : nullptr;
#pragma unroll
        for (int c8 = 0; c8 < TC_N / 8; ++c8) {
          uint4 hh, ll;
          __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(&hh);
          __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(&ll);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = v[c8 * 8 + i];
            hp[i] = __float2bfloat16_rn(x);
            lp[i] = __float2bfloat16_rn(x - __bfloat162float(hp[i]));
          }
          oh[c8] = hh;
          if (ol) ol[c8] = ll;
        }
      }
      if (a.headW) {   // fused 1x1 head: this thread holds all 64 channels of its row
        for (int d = 0; d < a.head_dim; ++d) {
          float s = __ldg(a.headB + d);
#pragma unroll
          for (int c = 0; c < TC_N; ++c) s = fmaf(v[c], __ldg(a.headW + c * a.head_dim + d), s);
          a.head_out[(size_t)grow * a.head_dim + d] = s;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// activations [B, L, C] bf16 -> box {64 channels, Lbox positions (element stride lstride), samples}
int tc_make_act_map(CUtensorMap* m, const void* base, int B, int L, int C, int box_l, int lstride, int box_b) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return B2P_ERR_NO_DEVICE;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)TC_K, (cuuint32_t)box_l, (cuuint32_t)box_b};
  cuuint32_t estr[3] = {1, (cuuint32_t)lstride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2P_OK : B2P_ERR_INVALID_ARG;
}
// weights [rows = taps*Cout, K = Cin] bf16 -> box {64 (K), 64 rows}
int tc_make_weight_map(CUtensorMap* m, const void* base, int rows, int K) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return B2P_ERR_NO_DEVICE;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_K, (cuuint32_t)TC_N};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2P_OK : B2P_ERR_INVALID_ARG;
}

template <int NSPLIT>
static int launch_t(const TcMaps& maps, const TcArgs& a, dim3 grid, cudaStream_t s) {
  constexpr int smem = NumStages<NSPLIT>::value * StageBytes<NSPLIT>::value + (int)sizeof(TcBarriers) + 1024;
  B2P_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  conv_tc_kernel<NSPLIT><<<grid, TC_THREADS, smem, s>>>(maps, a);
  return (int)cudaGetLastError();
}

int launch_conv_tc(const TcMaps& maps, const TcArgs& a, int nsplit, int nparity, cudaStream_t s) {
  if (a.Cout % TC_N || a.C[0] % TC_K || a.C[1] % TC_K || a.RC[0] % TC_K || a.RC[1] % TC_K || a.nrows <= 0) return B2P_ERR_INVALID_ARG;
  if (a.Lrows > 32 || (a.Lrows & (a.Lrows - 1)) || TC_M % a.Lrows) return B2P_ERR_INVALID_ARG;
  if (a.gn_gamma && (TC_N % a.cg != 0)) return B2P_ERR_INVALID_ARG;
  if (a.headW && a.Cout != TC_N) return B2P_ERR_INVALID_ARG;
  dim3 grid((a.nrows + TC_M - 1) / TC_M, a.Cout / TC_N, nparity);
  return nsplit == 2 ? launch_t<2>(maps, a, grid, s) : launch_t<1>(maps, a, grid, s);
}

}  // namespace b2p
