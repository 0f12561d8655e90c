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
// * epilogue on all 16 warps: warp w owns TMEM lane quadrant (w & 3) and the 16-column slice (w >> 2) of the 64-column
//   tile, so a thread holds 16 channels of one tile row: tap combine (row-shift shuffles), bias, GroupNorm(8) statistics
//   by warp shuffles over the L rows of a sample (+ a shared-memory exchange between column slices when a group is wider
//   than 16 channels), Mish, + time embedding, + residual, optional fused 1x1 head, bf16 hi/lo store.
// Thread 0 is the TMA producer and thread 32 the MMA issuer before they join the epilogue; warp 1 owns the TMEM allocation.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace b2p {

constexpr int TC_M = 128;           // rows per tile
constexpr int TC_N = 64;            // output channels per tile
constexpr int TC_K = 64;            // channels per pipeline stage (128 bytes of bf16: one swizzle atom row)
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 512;
constexpr int EPI_COLS = 16;          // columns per epilogue thread
constexpr int A_BYTES = TC_M * TC_K * 2;   // 16 KB
constexpr int B_BYTES = TC_N * TC_K * 2;   //  8 KB
constexpr int SLOT_BYTES = 16 * 1024;      // TMA ring slot


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
// long wait (epilogue warps waiting for the accumulators): suspend in hardware instead of spinning so that the waiting
// warps do not take issue slots from the single TMA / MMA issuing threads that share their schedulers
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITS_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONES_%=;\n\t"
      "bra WAITS_%=;\n\t"
      "DONES_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
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
// kind::f16: D fp32, A/B bf16, both K-major, M=128, N=n (multiple of 16, <= 256)
__device__ __forceinline__ constexpr uint32_t umma_idesc_n(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Mish with the SFU approximations (ex2.approx / rcp.approx): relative error ~1e-6, far below the bf16-split noise.
__device__ __forceinline__ float mish_fast(float x) {
  float e = __expf(x);
  float n = e * (e + 2.f);
  float m = x * __fdividef(n, n + 2.f);
  return x > 20.f ? x : m;
}

// GroupNorm(8) + Mish for the 16 channels a thread holds.  A group is CG consecutive channels x the L rows (adjacent
// lanes) of a sample.  CG <= 16: the group is local to the thread's slice; CG = 32 / 64: partial sums of the 2 / 4
// column-slice warps covering the group are exchanged through shared memory (xchg[row][slice]).  Two passes (mean, then
// centred variance) like the reference's GroupNorm.  Called by ALL threads of the CTA (contains __syncthreads).
template <int CG>
__device__ __forceinline__ void group_norm_mish16(float (&v)[EPI_COLS], int L, int row, int slice, float (*xchg)[4],
                                                  const float* gamma, const float* beta) {
  constexpr int W = CG < EPI_COLS ? CG : EPI_COLS;      // channels of one group inside this thread
  constexpr int NG = EPI_COLS / W;                      // groups per thread
  constexpr int SL = CG / W;                            // slices sharing one group
  const float inv_n = 1.0f / (float)(CG * L);
  float mean[NG], rstd[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < W; ++c) s += v[g * W + c];
    for (int o = 1; o < L; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean[g] = s;
  }
  if (SL > 1) {
    xchg[row][slice] = mean[0];
    __syncthreads();
    const int base = slice & ~(SL - 1);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < SL; ++j) s += xchg[row][base + j];
    mean[0] = s;
    __syncthreads();
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    mean[g] *= inv_n;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < W; ++c) { float d = v[g * W + c] - mean[g]; q = fmaf(d, d, q); }
    for (int o = 1; o < L; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd[g] = q;
  }
  if (SL > 1) {
    xchg[row][slice] = rstd[0];
    __syncthreads();
    const int base = slice & ~(SL - 1);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < SL; ++j) q += xchg[row][base + j];
    rstd[0] = q;
    __syncthreads();
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const float r = rsqrtf(rstd[g] * inv_n + 1e-5f);
#pragma unroll
    for (int c = 0; c < W; ++c) v[g * W + c] = mish_fast((v[g * W + c] - mean[g]) * r * gamma[g * W + c] + beta[g * W + c]);
  }
}

struct __align__(16) TcBarriers {
  uint64_t full[16];
  uint64_t empty[16];
  uint64_t tmem_full;
  uint32_t tmem_base;
  uint32_t pad;
  float bias[TC_N], gamma[TC_N], beta[TC_N], resb[TC_N];   // per-tile epilogue vectors
};
// epilogue scratch aliases the (by then idle) TMA ring:
//   red[2][KS][128/KS][64] fp32   split-K partial tiles received from the cluster peers        (64 KB)
//   xchg[128][4]                  GroupNorm partial sums between column slices                  ( 2 KB)
//   head[128][4][8]               fused-head partial dot products                               (16 KB)
constexpr int EPI_RED_BYTES = 2 * TC_M * TC_N * 4;
constexpr int EPI_XCHG_BYTES = TC_M * 4 * 4;

constexpr int TC_SMEM_STAGE_REGION = 224 * 1024;   // bytes available to the TMA ring
constexpr int TC_SMEM_TOTAL = TC_SMEM_STAGE_REGION + 1024 /*alignment slack*/ + (int)sizeof(TcBarriers);

// programmatic dependent launch: block until the preceding kernel in the stream has completed and flushed its writes /
// allow the next kernel in the stream to be scheduled (its pre-wait prologue then overlaps the rest of this kernel)
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f4(uint32_t local_saddr, uint32_t cta, float x, float y, float z, float w) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(cta));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// "taps in N": one pass over the activations computes Y_t = A * W_t^T for every tap t into its own 64-column TMEM
// block; the conv sum  out[l] = sum_t Y_t[l + shift_t]  is a ROW shift of the accumulator, done in the epilogue with warp
// shuffles (the L rows of a sample are adjacent lanes).  The activation tile is therefore loaded once per 64-channel
// chunk instead of once per tap, and zero padding is just "source lane outside the sample".
//
// Split-K over a thread-block cluster (gridDim.z == cluster size KS): the 64-channel K chunks of a layer are dealt
// round-robin to the KS CTAs of a cluster, so KS times more SMs stream the layer's weights/activations.  Each CTA
// combines its taps, then the partial [128 x 64] tiles are reduce-scattered BY ROWS through distributed shared memory:
// CTA j receives rows [j*128/KS, (j+1)*128/KS) from every peer, sums them in a fixed order (deterministic) and runs
// the GroupNorm/Mish/residual epilogue for those rows only (whole samples, so GroupNorm stays CTA-local).
template <int NSPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem + TC_SMEM_STAGE_REGION);

  const int T = a.T;
  const int stage_bytes = NSPLIT * (A_BYTES + T * B_BYTES);   // [A hi | A lo | W hi (T taps) | W lo (T taps)]
  int stages = TC_SMEM_STAGE_REGION / stage_bytes;
  if (stages > 8) stages = 8;
  const uint32_t tmem_cols = (T + 1) * TC_N <= 128 ? 128u : ((T + 1) * TC_N <= 256 ? 256u : 512u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, n0 = blockIdx.y * TC_N;
  const int KS = gridDim.z, rank = blockIdx.z;      // cluster = (1, 1, KS): rank == %cluster_ctarank
  const int b0 = tile_m * a.samples_per_tile;

  const int chunks0 = a.C[0] / TC_K, chunks1 = a.C[1] / TC_K;
  const int main_iters = chunks0 + chunks1;
  const int rchunks0 = a.RC[0] / TC_K, rchunks1 = a.RC[1] / TC_K;
  const int res_iters = rchunks0 + rchunks1;
  const int total_iters = (a.dbg & 1) ? 0 : main_iters + res_iters;
  // this CTA's share: global iterations rank, rank + KS, ...
  const int n_local = total_iters > rank ? (total_iters - rank + KS - 1) / KS : 0;
  const int n_main_local = (a.dbg & 1) ? 0 : (main_iters > rank ? (main_iters - rank + KS - 1) / KS : 0);
  const int n_res_local = n_local - n_main_local;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    mbar_init(&bars->tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + TC_N) {   // stage the tile's epilogue vectors
    const int c = threadIdx.x - 64;
    bars->bias[c] = a.bias ? __ldg(a.bias + n0 + c) : 0.f;
    bars->gamma[c] = a.gn_gamma ? __ldg(a.gn_gamma + n0 + c) : 1.f;
    bars->beta[c] = a.gn_gamma ? __ldg(a.gn_beta + n0 + c) : 0.f;
    bars->resb[c] = a.resB ? __ldg(a.resB + n0 + c) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (threadIdx.x == 0) {
    // =============================== TMA producer (one thread) ===============================
    prefetch_tmap(&maps.a[0][0]);
    prefetch_tmap(&maps.w[0]);
    if (NSPLIT == 2) { prefetch_tmap(&maps.a[0][1]); prefetch_tmap(&maps.w[1]); }
    // The weights do not depend on the previous layer: their TMA loads for the first ring pass are issued BEFORE the
    // grid-dependency wait and overlap the tail of the preceding kernel; only the activation loads wait for it.
    auto issue = [&](int j, bool do_w, bool do_a) {
      const int it = rank + j * KS;
      const int s = j % stages;
      uint8_t* st = smem + s * stage_bytes;
      uint8_t* sb = st + NSPLIT * A_BYTES;
      const bool res_phase = it >= main_iters;
      const int ch = res_phase ? it - main_iters : it;
      const int nch0 = res_phase ? rchunks0 : chunks0;
      const int src = ch < nch0 ? 0 : 1;
      const int c0 = (src == 0 ? ch : ch - nch0) * TC_K;
      const int kglob = (src == 0 ? 0 : (res_phase ? a.RC[0] : a.C[0])) + c0;
      if (do_w) {
        mbar_expect_tx(&bars->full[s], res_phase ? NSPLIT * (A_BYTES + B_BYTES) : stage_bytes);
#pragma unroll
        for (int h = 0; h < NSPLIT; ++h) {
          if (res_phase) tma_load_2d(sb + h * T * B_BYTES, &maps.rw[h], &bars->full[s], kglob, n0);
          else tma_load_3d(sb + h * T * B_BYTES, &maps.w[h], &bars->full[s], kglob, n0, a.tap0);
        }
      }
      if (do_a) {
#pragma unroll
        for (int h = 0; h < NSPLIT; ++h)
          tma_load_3d(st + h * A_BYTES, res_phase ? &maps.r[src][h] : &maps.a[src][h], &bars->full[s], c0, 0, b0);
      }
    };
    const int npre = n_local < stages ? n_local : stages;
    for (int j = 0; j < npre; ++j) issue(j, true, false);
    griddep_wait();
    for (int j = 0; j < npre; ++j) issue(j, false, true);
    for (int j = npre; j < n_local; ++j) {
      mbar_wait(&bars->empty[j % stages], ((j / stages) & 1) ^ 1);
      issue(j, true, true);
    }
  } else if (threadIdx.x == 32) {
    // =============================== MMA issuer (one thread) ===============================
    // All taps of a chunk are adjacent in shared memory ([T*64 rows] x 128 B, K-major), so ONE tcgen05.mma with
    // N = T*64 (<= 256; a fifth tap takes a second N = 64 instruction) covers them: the issuing thread stays far ahead
    // of the tensor pipe.  Descriptors are built once per stage and advanced by adding to the address field.
    const int nA = T < 4 ? T : 4;                                       // taps covered by the first instruction
    const uint32_t idescA = umma_idesc_n(nA * TC_N), idescB = umma_idesc_n(TC_N);
    bool main_first = true, res_first = true;
    for (int j = 0; j < n_local; ++j) {
      const int it = rank + j * KS;
      const int s = j % stages;
      const uint32_t ph = (j / stages) & 1;
      mbar_wait(&bars->full[s], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
      const uint32_t sb = sa + NSPLIT * A_BYTES;
      const bool res_phase = it >= main_iters;
      const bool first = res_phase ? res_first : main_first;
      if (res_phase) res_first = false; else main_first = false;
      const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + A_BYTES);
      const uint64_t b_hi = umma_desc(sb), b_lo = umma_desc(sb + T * B_BYTES);
      if (!res_phase) {
#pragma unroll
        for (int k = 0; k < TC_K / TC_UMMA_K; ++k) {
          const uint64_t ko = (uint64_t)(k * TC_UMMA_K * 2 / 16);       // +32 B per K step, in 16-byte units
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          umma(tmem_base, a_hi + ko, b_hi + ko, idescA, acc);
          if (NSPLIT == 2) {
            umma(tmem_base, a_lo + ko, b_hi + ko, idescA, 1u);
            umma(tmem_base, a_hi + ko, b_lo + ko, idescA, 1u);
          }
          if (T > 4) {
            const uint64_t t4 = (uint64_t)(4 * B_BYTES / 16);
            umma(tmem_base + 4 * TC_N, a_hi + ko, b_hi + t4 + ko, idescB, acc);
            if (NSPLIT == 2) {
              umma(tmem_base + 4 * TC_N, a_lo + ko, b_hi + t4 + ko, idescB, 1u);
              umma(tmem_base + 4 * TC_N, a_hi + ko, b_lo + t4 + ko, idescB, 1u);
            }
          }
        }
      } else {
        const uint32_t d = tmem_base + T * TC_N;
#pragma unroll
        for (int k = 0; k < TC_K / TC_UMMA_K; ++k) {
          const uint64_t ko = (uint64_t)(k * TC_UMMA_K * 2 / 16);
          umma(d, a_hi + ko, b_hi + ko, idescB, (first && k == 0) ? 0u : 1u);
          if (NSPLIT == 2) {
            umma(d, a_lo + ko, b_hi + ko, idescB, 1u);
            umma(d, a_hi + ko, b_lo + ko, idescB, 1u);
          }
        }
      }
      umma_commit(&bars->empty[s]);
    }
    umma_commit(&bars->tmem_full);
  }
  __syncwarp();
  griddep_wait();                 // everything below may read the previous kernels' outputs (residuals)
  griddep_launch_dependents();    // the next layer may start its prologue / weight prefetch on the idle SMs

  // =============================== epilogue (all 16 warps) ===============================
  {
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int slice = warp >> 2;                    // 16-column slice of the 64-column tile
    const int col0 = slice * EPI_COLS;
    const int r = quad * 32 + lane;                 // tile row == TMEM lane
    const int L = a.Lrows;
    const int rows_per_owner = TC_M / KS;
    const int r_local = r % rows_per_owner;
    const bool own = (r / rows_per_owner) == rank;  // this CTA finishes row r
    const long grow = (long)tile_m * TC_M + r;
    const bool row_ok = own && grow < a.nrows;
    const int b = (int)(grow >> a.log2L), l = (int)(grow & (L - 1));
    float* red = reinterpret_cast<float*>(smem);                                                  // [2][KS][rows_per_owner][64]
    float (*xchg)[4] = reinterpret_cast<float (*)[4]>(smem + EPI_RED_BYTES);
    float (*headp)[4][8] = reinterpret_cast<float (*)[4][8]>(smem + EPI_RED_BYTES + EPI_XCHG_BYTES);
    const int gcol = n0 + col0;
    const int n_out = (a.dbg & 2) ? 0 : a.n_out;
    // everything that is added AFTER GroupNorm/Mish (time embedding terms, identity residual, residual-conv bias) is
    // fetched from global memory now, while the tensor core is still busy
    float addv[EPI_COLS];
#pragma unroll
    for (int c = 0; c < EPI_COLS; ++c) addv[c] = res_iters > 0 ? bars->resb[col0 + c] : 0.f;
    if (row_ok) {
      if (a.temb) {
        const float* tp = a.temb + (size_t)b * a.temb_stride + gcol;
#pragma unroll
        for (int c = 0; c < EPI_COLS; c += 4) { float4 t4 = __ldg(reinterpret_cast<const float4*>(tp + c)); addv[c] += t4.x; addv[c + 1] += t4.y; addv[c + 2] += t4.z; addv[c + 3] += t4.w; }
      }
      if (a.temb2) {
#pragma unroll
        for (int c = 0; c < EPI_COLS; c += 4) { float4 t4 = __ldg(reinterpret_cast<const float4*>(a.temb2 + gcol + c)); addv[c] += t4.x; addv[c + 1] += t4.y; addv[c + 2] += t4.z; addv[c + 3] += t4.w; }
      }
      if (a.res_f32) {
        const float* q = a.res_f32 + (size_t)grow * a.Cout + gcol;
#pragma unroll
        for (int c = 0; c < EPI_COLS; c += 4) { float4 t4 = __ldg(reinterpret_cast<const float4*>(q + c)); addv[c] += t4.x; addv[c + 1] += t4.y; addv[c + 2] += t4.z; addv[c + 3] += t4.w; }
      }
      if (a.res_hi) {
        const uint4* qh = reinterpret_cast<const uint4*>(a.res_hi + (size_t)grow * a.Cout + gcol);
        const uint4* ql = a.res_lo ? reinterpret_cast<const uint4*>(a.res_lo + (size_t)grow * a.Cout + gcol) : nullptr;
#pragma unroll
        for (int c8 = 0; c8 < EPI_COLS / 8; ++c8) {
          uint4 hh = __ldg(qh + c8);
          const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&hh);
#pragma unroll
          for (int i = 0; i < 8; ++i) addv[c8 * 8 + i] += __bfloat162float(hp[i]);
          if (ql) {
            uint4 ll = __ldg(ql + c8);
            const __nv_bfloat16* lp = reinterpret_cast<const __nv_bfloat16*>(&ll);
#pragma unroll
            for (int i = 0; i < 8; ++i) addv[c8 * 8 + i] += __bfloat162float(lp[i]);
          }
        }
      }
    }
    mbar_wait_sleep(&bars->tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + col0;

    // ---- this CTA's partial results: slot o = output o (tap-combined), slot 1 = residual 1x1 block when n_out == 1 ----
    float part[2][EPI_COLS];
#pragma unroll
    for (int c = 0; c < EPI_COLS; ++c) { part[0][c] = 0.f; part[1][c] = 0.f; }
    if (n_main_local > 0) {
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        if (o >= n_out) break;
        for (int i = 0; i < a.nt[o]; ++i) {
          const int t = a.tap_blk[o][i], d = a.tap_shift[o][i];
          const bool valid = (l + d >= 0) && (l + d < L);
          const int src = (lane + d) & 31;
          float y[EPI_COLS];
          tmem_ld16(taddr + t * TC_N, y);
          if (d == 0) {
#pragma unroll
            for (int c = 0; c < EPI_COLS; ++c) part[o][c] += y[c];
          } else {
#pragma unroll
            for (int c = 0; c < EPI_COLS; ++c) { float g = __shfl_sync(0xffffffffu, y[c], src); part[o][c] += valid ? g : 0.f; }
          }
        }
      }
    }
    const bool has_res = res_iters > 0;
    if (has_res && n_res_local > 0) tmem_ld16(taddr + T * TC_N, part[1]);
    const int nslots = (n_out == 2 || has_res) ? 2 : 1;

    if (KS > 1 && !(a.dbg & 4)) {
      // ---- reduce-scatter by rows through distributed shared memory ----
      cluster_sync_all();                            // every peer's TMA ring is idle: its memory may be overwritten
      const uint32_t owner = (uint32_t)(r / rows_per_owner);
      for (int sl = 0; sl < nslots; ++sl) {
        const uint32_t dst = smem_u32(red + (((size_t)sl * KS + rank) * rows_per_owner + r_local) * TC_N + col0);
#pragma unroll
        for (int c = 0; c < EPI_COLS; c += 4)
          st_cluster_f4(dst + c * 4, owner, sl == 0 ? part[0][c] : part[1][c], sl == 0 ? part[0][c + 1] : part[1][c + 1],
                        sl == 0 ? part[0][c + 2] : part[1][c + 2], sl == 0 ? part[0][c + 3] : part[1][c + 3]);
      }
      cluster_sync_all();                            // all partial tiles have landed
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        if (sl >= nslots) break;
#pragma unroll
        for (int c = 0; c < EPI_COLS; ++c) part[sl][c] = 0.f;
        for (int srcc = 0; srcc < KS; ++srcc) {      // fixed summation order => deterministic
          const float4* q = reinterpret_cast<const float4*>(red + (((size_t)sl * KS + srcc) * rows_per_owner + r_local) * TC_N + col0);
#pragma unroll
          for (int c4 = 0; c4 < EPI_COLS / 4; ++c4) { float4 t4 = q[c4]; part[sl][c4 * 4] += t4.x; part[sl][c4 * 4 + 1] += t4.y; part[sl][c4 * 4 + 2] += t4.z; part[sl][c4 * 4 + 3] += t4.w; }
        }
      }
      __syncthreads();                               // red[] fully consumed before xchg/head scratch (disjoint) — keeps phases tidy
    }

#pragma unroll
    for (int o = 0; o < 2; ++o) {
      if (o >= n_out) break;
      float v[EPI_COLS];
#pragma unroll
      for (int c = 0; c < EPI_COLS; ++c) v[c] = part[o][c] + bars->bias[col0 + c];
      if (a.gn_gamma) {
        switch (a.cg) {
          case 8: group_norm_mish16<8>(v, L, r, slice, xchg, bars->gamma + col0, bars->beta + col0); break;
          case 16: group_norm_mish16<16>(v, L, r, slice, xchg, bars->gamma + col0, bars->beta + col0); break;
          case 32: group_norm_mish16<32>(v, L, r, slice, xchg, bars->gamma + col0, bars->beta + col0); break;
          default: group_norm_mish16<64>(v, L, r, slice, xchg, bars->gamma + col0, bars->beta + col0); break;
        }
      }
      const bool ok = row_ok && (a.out_ldiv == 1 || (l % a.out_ldiv) == 0);
#pragma unroll
      for (int c = 0; c < EPI_COLS; ++c) v[c] += addv[c];
      if (has_res) {
#pragma unroll
        for (int c = 0; c < EPI_COLS; ++c) v[c] += part[1][c];
      }
      if (ok) {
        const size_t orow = (size_t)b * a.out_L + (size_t)(l / a.out_ldiv) * a.out_lmul + o;
        if (a.out_hi) {
          uint4* oh = reinterpret_cast<uint4*>(a.out_hi + orow * a.Cout + gcol);
          uint4* ol = a.out_lo ? reinterpret_cast<uint4*>(a.out_lo + orow * a.Cout + gcol) : nullptr;
#pragma unroll
          for (int c8 = 0; c8 < EPI_COLS / 8; ++c8) {
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
      }
      if (a.headW) {   // fused 1x1 head (Cout == 64): partial dot products per column slice, summed by slice 0
        for (int d = 0; d < a.head_dim; ++d) {
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < EPI_COLS; ++c) s = fmaf(v[c], __ldg(a.headW + (col0 + c) * a.head_dim + d), s);
          headp[r][slice][d] = s;
        }
        __syncthreads();
        if (slice == 0 && ok) {
          for (int d = 0; d < a.head_dim; ++d)
            a.head_out[(size_t)grow * a.head_dim + d] = __ldg(a.headB + d) + headp[r][0][d] + headp[r][1][d] + headp[r][2][d] + headp[r][3][d];
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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
// conv weights [taps][Cout][Cin] bf16 -> box {64 (K), 64 output channels, ntaps}; ntaps == 0 -> plain 2-D [Cout][Cin] map
int tc_make_weight_map(CUtensorMap* m, const void* base, int taps, int Cout, int K, int box_taps) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return B2P_ERR_NO_DEVICE;
  CUresult r;
  if (box_taps == 0) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_K, (cuuint32_t)TC_N};
    cuuint32_t estr[2] = {1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Cout, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)Cout * K * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_K, (cuuint32_t)TC_N, (cuuint32_t)box_taps};
    cuuint32_t estr[3] = {1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? B2P_OK : B2P_ERR_INVALID_ARG;
}

template <int NSPLIT>
static int launch_t(const TcMaps& maps, const TcArgs& a, dim3 grid, cudaStream_t s) {
  constexpr int smem = TC_SMEM_TOTAL;
  B2P_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = grid.z;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("B2P_TC_PDL"); pdl = e ? atoi(e) : 1; }
  cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
  return (int)cudaLaunchKernelEx(&cfg, conv_tc_kernel<NSPLIT>, maps, a);
}

// split-K factor: as many cluster CTAs as (a) there are K chunks, (b) keeps whole samples per owner (128/KS >= L),
// (c) fits the launch in roughly one wave of the 148 SMs
static int pick_ksplit(const TcArgs& a, int tiles) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("B2P_TC_KSPLIT"); forced = e ? atoi(e) : 0; }
  const int iters = (a.C[0] + a.C[1] + a.RC[0] + a.RC[1]) / TC_K;
  int ks = 1;   // split-K is opt-in (B2P_TC_KSPLIT): the DSMEM reduce-scatter costs more than it saves at these sizes
  (void)tiles;
  if (forced > 0) { ks = 1; while (ks < forced && ks * 2 <= iters && TC_M / (ks * 2) >= a.Lrows) ks *= 2; }
  return ks;
}

int launch_conv_tc(const TcMaps& maps, const TcArgs& a_in, int nsplit, cudaStream_t s) {
  TcArgs a = a_in;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("B2P_TC_DBG"); dbg = e ? atoi(e) : 0; } a.dbg = dbg; }
  if (a.T < 1 || a.T > 5 || a.n_out < 1 || a.n_out > 2 || a.out_ldiv < 1) return B2P_ERR_INVALID_ARG;
  if (a.Cout % TC_N || a.C[0] % TC_K || a.C[1] % TC_K || a.RC[0] % TC_K || a.RC[1] % TC_K || a.nrows <= 0) return B2P_ERR_INVALID_ARG;
  if (a.Lrows > 32 || (a.Lrows & (a.Lrows - 1)) || TC_M % a.Lrows) return B2P_ERR_INVALID_ARG;
  if (a.gn_gamma && (TC_N % a.cg != 0)) return B2P_ERR_INVALID_ARG;
  if (a.headW && (a.Cout != TC_N || a.head_dim > 8)) return B2P_ERR_INVALID_ARG;
  if (a.n_out == 2 && (a.RC[0] || a.RC[1])) return B2P_ERR_INVALID_ARG;
  const int mt = (a.nrows + TC_M - 1) / TC_M, nt = a.Cout / TC_N;
  dim3 grid(mt, nt, pick_ksplit(a, mt * nt));
  return nsplit == 2 ? launch_t<2>(maps, a, grid, s) : launch_t<1>(maps, a, grid, s);
}

}  // namespace b2p
