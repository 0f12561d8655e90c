// Row-owned chain kernel for the 64-channel layers of the denoiser (K2 + K4 + K6 + K1 fused).
//
// At the full-resolution end of the U-Net every layer has 64 output channels, so ONE CTA can own a group of 8 trajectories
// (8 x 16 positions = one 128-row tcgen05 tile, all 64 channels = one column tile) through a whole CHAIN of layers without
// ever exchanging data with another CTA: activations stay in shared memory as the next layer's A operand (bf16 hi/lo,
// K-major, 128-byte swizzle — written by the epilogue in exactly the layout TMA would produce), accumulators in TMEM, weights
// stream in as pre-swizzled images (one bulk copy per tap, issued as soon as the previous layer's MMAs have retired).
// No kernel boundary, no dependency release, no first-operand latency between the layers of a chain.
// While the doubled grid still fits the SMs in one wave the group is owned by a CLUSTER of two CTAs instead (template parameter CL; four is built too):
// each computes one channel half of every layer and stores its chunks of the next A operand into both CTAs' shared memory, because the epilogue
// (tap combine + GroupNorm + Mish for 128 rows x 64 channels), not the tensor core, is what an op costs.  All forms give bit-identical results.
//
// Chains (modeling/temporal.py:197-245, interact.py:132-164):
//   head of an evaluation  : downs.0 = im2col'd Conv1dBlock(7->64) [+temb], Conv1dBlock(64->64) [+1x1 projection of x],
//                            second residual block, Downsample1d(64)                              (5 tensor-core layers)
//   tail of an evaluation  : ups.<last> from its second conv on, Upsample1d(64), final_conv / act_conv with the fused 1x1 head
//   seam (no guidance)     : tail of evaluation i + the scheduler step (bit-exact arithmetic of sched_math.cuh) + head of
//                            evaluation i+1 in ONE launch: x_t never leaves shared memory between two denoiser evaluations.
// Precision: NSPLIT = 2 -> bf16 hi/lo operands, hi*hi + lo*hi + hi*lo accumulated in fp32 (same "bf16x3" as conv_tc.cu);
// NSPLIT = 1 -> single bf16 pass.  GroupNorm / Mish / residuals / scheduler arithmetic in fp32.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <string.h>

#include "chain64.cuh"
#include "tc_ptx.cuh"

namespace b2p {

constexpr int CH_EPI_THREADS = 512;                  // 16 epilogue warps: 4 per TMEM lane quadrant
constexpr int CH_MMA_WARP = 16;                      // a 17th warp issues the MMAs AND, the moment they have retired, the next op's weight copy: an epilogue warp
                                                     // got to that copy only after its own half-0 epilogue, 1.5-2.5 k cycles later (the copy itself lands 1.4-2.0 k cycles after issue)
constexpr int CH_THREADS = CH_EPI_THREADS + 32;
constexpr int CH_EC = 8;                             // columns per epilogue thread and column half: one GroupNorm group, one 16-byte chunk
constexpr int CH_SL = 4;                             // column slices per half (4 warps per TMEM lane quadrant)
constexpr int CH_HCOLS = 160;                        // TMEM column stride between the two column halves (5 taps x 32 channels)
constexpr int CH_HALF = 16 * 1024;                    // one bf16 plane of an activation buffer: [128 rows][64 ch]
constexpr int CH_ACT_BYTES = 2 * CH_HALF;             // hi | lo
constexpr int CH_TAP_BYTES = 64 * 64 * 2;             // one tap of one weight plane
constexpr int CH_OFF_W = 3 * CH_ACT_BYTES;            // weight image of the current op (<= 5 taps x 8 KB x 2 planes)
constexpr int CH_W_BYTES = 2 * 5 * CH_TAP_BYTES;
constexpr int CH_OFF_MISC = CH_OFF_W + CH_W_BYTES;
constexpr int CH_HEADP_SLOTS = 2 * CH_SL;             // head partial sums [column half x slice][8][128 rows] = 32 KB: they live in an activation buffer that is idle
static_assert(CH_HEADP_SLOTS * 8 * 128 * 4 <= CH_ACT_BYTES, "head partial sums alias one activation buffer");   // during the head op (neither its A operand nor the next op's)
static_assert(CH_EPI_THREADS == 512, "the epilogue mapping assumes 16 warps: 4 lane quadrants x 4 column slices of 8 channels per half");
#ifndef CH_STASYNC
#define CH_STASYNC 0   // cluster forms, compile-time experiment: 1 = the next A operand travels as st.async stores that complete a byte count on the CONSUMER's
#endif                 // mbarrier (no writer-side drain, no cluster barrier per op).  Built, bit-identical, and 0.5 % SLOWER per plan than 0 = st.shared::cluster +
                       // barrier.cluster release / acquire per op (same-call A/B at B = 256: 241.4 vs 240.2 us per iteration): the STAS stores lengthen the pass
constexpr int CH_MAX_HD = 16 * 8;                     // horizon 16 x transition dim <= 8 (im2col K = 5 * D <= 64; one scheduler element per thread)

struct __align__(16) ChainShared {
  uint64_t wbar;                 // weight image of the current op has landed
  uint64_t mma_bar[2];           // MMAs of column half 0 / 1 of the current op have retired
  uint64_t abar[2];              // cluster forms: the peers' chunks of op oi's A operand have landed (by op parity)
  uint32_t tmem_base;
  uint32_t pad;                  // the float tables below are read as float4: keep them 16-byte aligned
  float vec[CH_MAXOPS][3][64];   // bias / gamma / beta of every op, fetched once in the prologue (a per-op fetch put one global-load latency on every op-to-op hand-over)
  float xs[CH_NS * CH_MAX_HD];   // x_t of the CTA's trajectories
  float mo[CH_NS * CH_MAX_HD];   // model output of the CTA's trajectories (head result)
  float xw[8 * 64 + 64];         // 1x1 projection of x: [D][64] + bias
  float headw[8 * 64 + 8];       // head weights [d][64] + bias
  SchedK sk;                     // the fused scheduler step's arguments: the out-of-line step function takes them by reference, and a reference into the kernel
                                 // parameters is a GENERIC pointer whose every field read is a global-path round trip (9 k cycles per launch)
  ChainOp ops[CH_MAXOPS];        // the op table, copied out of the kernel-parameter constant bank in the prologue: an indexed read of a.ops[oi]
                                 // is a constant-cache miss the first time an op's line is touched, and those misses sat on the op-to-op critical path
};

static_assert(CH_OFF_MISC + sizeof(ChainShared) + 1024 <= 227 * 1024, "shared memory budget");
static_assert(offsetof(ChainShared, xs) % 16 == 0 && offsetof(ChainShared, mo) % 16 == 0 && offsetof(ChainShared, headw) % 16 == 0, "float4 alignment");

size_t chain64_smem_bytes() { return CH_OFF_MISC + sizeof(ChainShared) + 1024; }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of the 16-byte chunk `c8` (8 channels) of row `r` inside a K-major, 128-byte-swizzled [128][64] bf16 plane
__device__ __forceinline__ uint32_t swz(int r, int c8) { return (uint32_t)(r * 128 + ((c8 ^ (r & 7)) << 4)); }

__device__ __forceinline__ void split8(const float* v, uint4* hi, uint4* lo) {
  __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(hi);
  __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(lo);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    hp[i] = __float2bfloat16_rn(v[i]);
    lp[i] = __float2bfloat16_rn(v[i] - __bfloat162float(hp[i]));
  }
}

// The scheduler step runs once per launch: kept out of line so the hot per-op loop body stays small in the instruction cache.
__device__ __noinline__ float sched_elem(const SchedK& k, int sample, int pos, int col, float m, float x, float nz, float tj, float mk, float* x0) {
  return step_one(k, sample, pos, col, m, 0.f, x, nz, tj, mk, x0);
}
__device__ __noinline__ float4 philox_group(unsigned long long seed, unsigned group, unsigned step) { return philox_normal4(seed, group, step); }

// ---- epilogue building blocks with COMPILE-TIME tap counts / shifts / group sizes: the shuffles sit in straight-line code, so the
// compiler emits plain SHFL instead of wrapping every one in a convergence barrier (the epilogue is instruction-issue bound) ----
// v[c] += sum_i Y_blk(i)[l + d(i)] (zero outside the trajectory)
template <int NT>
__device__ __forceinline__ void tap_combine(float (&v)[CH_EC], uint32_t tbase, const int (&blk)[NT], const int (&sh)[NT], int l, int L, int lane) {
  float y[NT][CH_EC];
#pragma unroll
  for (int i = 0; i < NT; ++i) tmem_ld<CH_EC, false>(tbase + blk[i] * 16, y[i]);   // every tap block in flight at once: ONE exposed TMEM latency per half
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int d = sh[i];
    if (d == 0) {
#pragma unroll
      for (int c = 0; c < CH_EC; ++c) v[c] += y[i][c];
    } else {
      const bool valid = (l + d >= 0) && (l + d < L);
      const int src = (lane + d) & 31;
#pragma unroll
      for (int c = 0; c < CH_EC; ++c) { const float g = __shfl_sync(0xffffffffu, y[i][c], src); if (valid) v[c] += g; }
    }
  }
}
// GroupNorm(8) + Mish of the thread's 8 channels: a group = these channels x the L rows (adjacent lanes) of a trajectory
template <int L>
__device__ __forceinline__ void group_norm8_mish(float (&v)[CH_EC], const float* gamma, const float* beta) {
  const float inv_n = 1.0f / (float)(8 * L);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += v[c];
#pragma unroll
  for (int of = 1; of < L; of <<= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
  const float mean = s * inv_n;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) { const float dd = v[c] - mean; q = fmaf(dd, dd, q); }
#pragma unroll
  for (int of = 1; of < L; of <<= 1) q += __shfl_xor_sync(0xffffffffu, q, of);
  const float rs = rsqrtf(q * inv_n + 1e-5f);
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = mish_fast((v[c] - mean) * rs * gamma[c] + beta[c]);
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) { uint32_t ra; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(saddr), "r"(cta)); return ra; }
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_cluster_1f(uint32_t raddr, float x) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(x) : "memory"); }

// CL = 1: one CTA owns 8 trajectories and all 64 channels (two column-half passes per op).
// CL = 2: a cluster of two CTAs owns the 8 trajectories; CTA `rank` computes channel half `rank` of every op (half the MMAs, ONE epilogue pass)
//         and stores its 16-byte chunks of the next A operand into BOTH CTAs' shared memory (st.shared::cluster); the op-to-op barrier is the
//         cluster barrier.  The epilogue is instruction-issue bound (4 warps per scheduler x ~620 instructions per pass), so splitting the
//         channels over two SMs is what shortens an op; the arithmetic per element is identical, results are bit-identical to CL = 1.
template <int NSPLIT, int CL>
__global__ void __launch_bounds__(CH_THREADS, 1) chain64_kernel(const __grid_constant__ ChainArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  ChainShared* sh = reinterpret_cast<ChainShared*>(smem + CH_OFF_MISC);
  uint8_t* wbuf = smem + CH_OFF_W;
  constexpr int NH = CL == 1 ? 2 : 1;                   // column-half passes per CTA
  constexpr int NPEER = CL - 1;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  uint32_t peer_base[NPEER > 0 ? NPEER : 1];            // the other CTAs' copies of `smem`
#pragma unroll
  for (int pr = 0; pr < NPEER; ++pr) peer_base[pr] = mapa_u32(smem_u32(smem), (uint32_t)(crank ^ (pr + 1)));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = (int)(blockIdx.x / CL) * a.ns;
  const int nb = a.B - b0 < a.ns ? a.B - b0 : a.ns;
  const int HD = a.H * a.D;

  if (tid == 0) {
    mbar_init(&sh->wbar, 1);
    mbar_init(&sh->mma_bar[0], 1);
    mbar_init(&sh->mma_bar[1], 1);
    mbar_init(&sh->abar[0], 1);
    mbar_init(&sh->abar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == CH_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows that no trajectory of this CTA owns are never written: start from zero so the tensor core never reads stale bit patterns
  for (int i = tid; i < 3 * CH_ACT_BYTES / 16; i += CH_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  // weights-only tables: independent of the preceding kernels
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.ops);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sh->ops);
    for (int i = tid; i < (int)(a.n_ops * sizeof(ChainOp) / 4); i += CH_THREADS) dst[i] = src[i];
  }
  if (a.do_sched) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&a.sk);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh->sk);
    for (int i = tid; i < (int)(sizeof(SchedK) / 4); i += CH_THREADS) dst[i] = src[i];
  }
  for (int i = tid; i < a.n_ops * 192; i += CH_THREADS) {
    const int o = i / 192, w = (i >> 6) % 3, c = i & 63;
    const float* p = w == 0 ? a.ops[o].bias : (w == 1 ? a.ops[o].gamma : a.ops[o].beta);
    sh->vec[o][w][c] = p ? __ldg(p + c) : (w == 1 ? 1.f : 0.f);
  }
  if (a.xprojW)
    for (int i = tid; i < a.D * 64 + 64; i += CH_THREADS) sh->xw[i] = i < a.D * 64 ? __ldg(a.xprojW + i) : __ldg(a.xprojB + i - a.D * 64);
  if (a.headW) {   // [64][head_dim] -> [d][64]
    const int hd = a.head_dim;
    for (int i = tid; i < 64 * hd; i += CH_THREADS) { const int c = i / hd, d = i - c * hd; sh->headw[d * 64 + c] = __ldg(a.headW + i); }
    if (tid < hd) sh->headw[8 * 64 + tid] = __ldg(a.headB + tid);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // every CTA of the cluster has initialised its barriers and zeroed its buffers before a peer stores into them
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  auto issue_weights = [&](int oi) {   // one thread: pre-swizzled image, one bulk copy per tap and plane
    const ChainOp& op = sh->ops[oi];
    const uint32_t plane = (uint32_t)op.T * CH_TAP_BYTES;
    const uint8_t* src = a.wpack + op.w_off;
    if (CL > 1) {                        // only this CTA's channel half / quarter of each plane (the image is ordered [plane][channel quarter][tap][16 channels]), same offsets
      const uint32_t part_bytes = plane / CL, o0 = (uint32_t)crank * part_bytes;
      mbar_expect_tx(&sh->wbar, NSPLIT * part_bytes);
      for (int pl = 0; pl < NSPLIT; ++pl)
        for (int i = 0; i < op.T; ++i) {
          const uint32_t o = pl * plane + o0 + (uint32_t)i * (CH_TAP_BYTES / CL);
          bulk_g2s(wbuf + o, src + o, CH_TAP_BYTES / CL, &sh->wbar);
        }
    } else {
      mbar_expect_tx(&sh->wbar, NSPLIT * plane);
      for (int i = 0; i < NSPLIT * op.T; ++i) bulk_g2s(wbuf + (size_t)i * CH_TAP_BYTES, src + (size_t)i * CH_TAP_BYTES, CH_TAP_BYTES, &sh->wbar);
    }
  };
  if (warp == CH_MMA_WARP && elect_one()) issue_weights(0);
  griddep_launch_dependents();      // the next kernel may become resident and prefetch ITS weights
  griddep_wait();                   // everything below reads what the preceding kernels wrote

  // x_t of this CTA's trajectories
  for (int i = tid; i < CH_NS * HD; i += CH_THREADS) {
    const int s = i / HD;
    float v = 0.f;
    if (s < nb) {
      const int bs = b0 + s;
      const int row = a.x_period > 0 ? bs % a.x_period : bs;
      v = a.x[(size_t)row * HD + (i - s * HD)];   // plain load: a seam launch rewrites x in place later on
    }
    sh->xs[i] = v;
  }
  if (a.ops[0].in_buf >= 0) {       // the chain starts from a tensor in global memory: stage it as the first A operand
    const int L0 = a.ops[0].L;
    uint8_t* dst = smem + a.ops[0].in_buf * CH_ACT_BYTES;
    const int rows_ok = nb * L0;
    for (int i = tid; i < NSPLIT * TC_M * 8; i += CH_THREADS) {
      const int half = i >> 10, row = (i >> 3) & 127, c8 = i & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < rows_ok) v = __ldg(reinterpret_cast<const uint4*>((half ? a.in_lo : a.in_hi) + ((size_t)b0 * L0 + row) * 64 + c8 * 8));
      *reinterpret_cast<uint4*>(dst + half * CH_HALF + swz(row, c8)) = v;
    }
  }

  const int quad = warp & 3, slice = warp >> 2;
  const int r = quad * 32 + lane;                               // tile row == TMEM lane
  const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
  // this warp's 8-channel chunk (of the 64 channels) in column-half pass hf, and whether the warp has one at all (a quarter is two chunks: slices 0, 1)
  const bool wactive = warp < CH_EPI_THREADS / 32 && (CL < 4 || slice < 2);
  auto chunk_of = [&](int hf) { return CL == 1 ? hf * 4 + slice : (CL == 2 ? crank * 4 + slice : crank * 2 + (slice & 1)); };
  uint32_t wpar = 0, mpar = 0;
  uint32_t aph = 0;                 // phase bits of abar[0], abar[1]: a phase completes only on an op that waits for peers' chunks (MMA lane only)

#pragma unroll 1
  for (int oi = 0; oi < a.n_ops; ++oi) {
    const ChainOp& op = sh->ops[oi];
    const int L = op.L;
    if (op.in_buf == CH_IN_IM2COL) {
      // A[row (s, l)][k = j * D + c] = x[s, l + j - pad, c]  (zero outside the trajectory, zero for k >= T_im2col * D)
      __syncthreads();                                           // xs may just have been rewritten by the scheduler step
      const int D = a.D, KT = 5 * D;
      for (int i = tid; i < TC_M * 8; i += CH_THREADS) {
        const int row = i >> 3, c8 = i & 7;
        const int s = row >> 4, l = row & 15;                    // H == 16
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // k = j * D + c and pos = l + j - 2, so the element is xs[s][(l - 2) * D + k]; it lies inside the trajectory exactly when that offset does
          const int k = c8 * 8 + e, off = (l - 2) * D + k;
          v[e] = (k < KT && s < nb && off >= 0 && off < HD) ? sh->xs[s * HD + off] : 0.f;
        }
        uint4 hi, lo;
        split8(v, &hi, &lo);
        *reinterpret_cast<uint4*>(smem + swz(row, c8)) = hi;
        if (NSPLIT == 2) *reinterpret_cast<uint4*>(smem + CH_HALF + swz(row, c8)) = lo;
      }
    }
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[oi * 16 + 0] = clock64();
    // the A operand was written with ordinary stores: make the LOCAL ones visible to the tensor core here; the remote ones become visible to the peer through the
    // cluster barrier's release / acquire, and the peer's MMA lane runs its own proxy fence behind the barrier (a full `fence.proxy.async` here costs a second
    // GPU-scope membar on top of the one inside barrier.cluster.arrive.release)
    fence_async_smem();
    tc_fence_before();               // ... and the previous op's TMEM reads are complete
    // both CTAs have written their channel half of this op's A operand into both copies: ARRIVE now, wait after the address work and load issue below
    if (CL > 1 && !CH_STASYNC) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    // ------------- everything added after GroupNorm / Mish comes from global memory and depends on nothing this kernel computes: the loads are
    // ISSUED here, ahead of the barrier, all of them before the first use (issued after the barrier and consumed half by half they were two to
    // three serialised L2 round trips, 2.3-3.2 k cycles against the ~1 k of the first column half's MMAs) -------------
    const int l = r & (L - 1), sidx = r >> op.log2L;
    const bool epi = wactive;                                // the extra warp only issues MMAs; with four CTAs per group half of the epilogue warps have no chunk
    const bool row_ok = epi && sidx < nb;
    const int b = b0 + sidx;
    const bool has_t = row_ok && op.temb_off >= 0, has_q = row_ok && op.res_kind == CH_RES_F32;
    const float* t2p = a.temb2[op.phase];
    const bool has_t2 = has_t && t2p != nullptr;
    float4 raw[NH][3][CH_EC / 4];                            // [column-half pass][time row | per-step time vector | fp32 residual]
#pragma unroll
    for (int hf = 0; hf < NH; ++hf) {
      const int ch0 = chunk_of(hf) * 8;
#pragma unroll
      for (int j = 0; j < CH_EC / 4; ++j) {
        raw[hf][0][j] = raw[hf][1][j] = raw[hf][2][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_t) raw[hf][0][j] = __ldg(reinterpret_cast<const float4*>(a.temb_rows + (size_t)b * a.temb_stride + op.temb_off + ch0) + j);
        if (has_t2) raw[hf][1][j] = __ldg(reinterpret_cast<const float4*>(t2p + op.temb_off + ch0) + j);
        if (has_q) raw[hf][2][j] = __ldg(reinterpret_cast<const float4*>(a.res_f32 + ((size_t)b * L + l) * 64 + ch0) + j);
      }
    }
    if (CL > 1 && !CH_STASYNC) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();            // (st.async form: the CTA barrier covers this CTA's own chunks and its TMEM reads; the peers' chunks are counted on abar)
    if (a.trace && blockIdx.x == 0) {   // BAR.SYNC lets the next instruction issue before the warp blocks: a clock read right behind it is the ARRIVAL time
      __syncwarp();
      if (tid == 0) a.trace[oi * 16 + 1] = clock64();
    }

    if (warp == CH_MMA_WARP) {
      // =============================== MMA issue (one elected lane of the 17th warp runs the whole loop: tc_ptx.cuh elect_one) ===============================
      if (elect_one()) {
      if (a.trace && blockIdx.x == 0) a.trace[oi * 16 + 14] = clock64();
      mbar_wait(&sh->wbar, wpar);
      if (CL > 1 && CH_STASYNC && oi > 0 && op.in_buf >= 0) {
        // the peers' chunks of this op's A operand: every peer thread that stored sent 16 bytes per plane (and per output row of an upsampling op);
        // the peers run the same op table on the same trajectories, so the count is the one this CTA's own epilogue warps produced for them
        const ChainOp& pv = sh->ops[oi - 1];
        int nq = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) nq += (((q * 32) >> pv.log2L) < nb && (pv.kind != CH_UP || q < 2)) ? 1 : 0;
        const uint32_t per_cta = (uint32_t)nq * (CL == 4 ? 2u : 4u) * 32u * 16u * NSPLIT * (pv.kind == CH_UP ? 2u : 1u);
        mbar_expect_tx(&sh->abar[oi & 1], (CL - 1) * per_cta);
        mbar_wait(&sh->abar[oi & 1], (aph >> (oi & 1)) & 1u);
        aph ^= 1u << (oi & 1);
      }
      if (CL > 1) fence_async_smem();   // reader side of the stores the peer CTAs made into this CTA's A operand
      tc_fence_after();
      if (a.trace && blockIdx.x == 0) a.trace[oi * 16 + 2] = clock64();
      const uint32_t sa = smem_u32(smem + (op.in_buf < 0 ? 0 : op.in_buf) * CH_ACT_BYTES);
      const uint32_t sb = smem_u32(wbuf);
      const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + CH_HALF);
      // the weight image is ordered [column half][tap][32 channels]: the taps of one half are ONE instruction of N = 32 * T, and the
      // epilogue of half 0 overlaps the MMAs of half 1
      const uint32_t idn = umma_idesc_n(CL == 4 ? op.T * 16 : op.T * 32);
      const uint32_t half_bytes = (uint32_t)op.T * (CH_TAP_BYTES / (CL == 4 ? 4 : 2));   // one MMA operand: a channel half (a quarter with four CTAs)
#pragma unroll 1
      for (int hf = 0; hf < NH; ++hf) {
        const uint32_t hc = CL > 1 ? (uint32_t)crank : (uint32_t)hf;    // which half / quarter
        const uint64_t b_hi = umma_desc(sb + hc * half_bytes), b_lo = umma_desc(sb + op.T * CH_TAP_BYTES + hc * half_bytes);
        const uint32_t dcol = tmem_base + hf * CH_HCOLS;
#pragma unroll 1
        for (int k = 0; k < op.ksteps; ++k) {
          const uint64_t ko = (uint64_t)(k * TC_UMMA_K * 2 / 16);  // +32 B per K step, in 16-byte units
          umma(dcol, a_hi + ko, b_hi + ko, idn, k == 0 ? 0u : 1u);
          if (NSPLIT == 2) {
            umma(dcol, a_lo + ko, b_hi + ko, idn, 1u);
            umma(dcol, a_hi + ko, b_lo + ko, idn, 1u);
          }
        }
        umma_commit(&sh->mma_bar[hf]);
      }
      if (a.trace && blockIdx.x == 0) a.trace[oi * 16 + 3] = clock64();
      if (oi + 1 < a.n_ops) {          // the weight buffer is free once every MMA of this op has retired
        mbar_wait(&sh->mma_bar[NH - 1], mpar);
        if (a.trace && blockIdx.x == 0) a.trace[(oi + 1) * 16 + 12] = clock64();
        issue_weights(oi + 1);
      }
      }
      __syncwarp();
    }
    wpar ^= 1;
    __syncwarp();

    // ------------- the addends, summed in the order the per-layer kernels use (time row, per-step time vector, residual) -------------
    float addv[NH][CH_EC];
#pragma unroll
    for (int hf = 0; hf < NH; ++hf) {
      const int ch0 = chunk_of(hf) * 8;                     // first of this thread's 8 channels in this pass
#pragma unroll
      for (int c = 0; c < CH_EC; ++c) addv[hf][c] = 0.f;
      if (has_t) {
#pragma unroll
        for (int j = 0; j < CH_EC / 4; ++j) {
          const float4 t4v = raw[hf][0][j];
          addv[hf][4 * j] += t4v.x; addv[hf][4 * j + 1] += t4v.y; addv[hf][4 * j + 2] += t4v.z; addv[hf][4 * j + 3] += t4v.w;
          if (has_t2) { const float4 u = raw[hf][1][j]; addv[hf][4 * j] += u.x; addv[hf][4 * j + 1] += u.y; addv[hf][4 * j + 2] += u.z; addv[hf][4 * j + 3] += u.w; }
        }
      }
      if (has_q) {
#pragma unroll
        for (int j = 0; j < CH_EC / 4; ++j) {
          const float4 t4v = raw[hf][2][j];
          addv[hf][4 * j] += t4v.x; addv[hf][4 * j + 1] += t4v.y; addv[hf][4 * j + 2] += t4v.z; addv[hf][4 * j + 3] += t4v.w;
        }
      } else if (row_ok && op.res_kind == CH_RES_XPROJ) {   // residual_conv(x) of the first block: 1x1 conv over the D raw channels, fp32 on CUDA cores
        const float* xr = sh->xs + sidx * HD + l * a.D;
#pragma unroll
        for (int c = 0; c < CH_EC; ++c) addv[hf][c] = sh->xw[a.D * 64 + ch0 + c];
#pragma unroll 4
        for (int d = 0; d < 8; ++d) {                       // D <= 8: straight-line code, the shared-memory reads of several d in flight together
          if (d < a.D) {
            const float xv = xr[d];
#pragma unroll
            for (int c = 0; c < CH_EC; ++c) addv[hf][c] = fmaf(xv, sh->xw[d * 64 + ch0 + c], addv[hf][c]);
          }
        }
      }
    }
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[oi * 16 + 4] = clock64();

    // =============================== epilogue (all 16 warps), one column half at a time ===============================
    const int n_out = op.kind == CH_UP ? 2 : 1;
    const int hb = op.in_buf == 2 ? 0 : 2;              // fused head: the partial sums go to an activation buffer that is idle during this op
    float (*headp)[8][TC_M] = reinterpret_cast<float (*)[8][TC_M]>(smem + hb * CH_ACT_BYTES);
#pragma unroll
    for (int hf = 0; hf < NH; ++hf) {
      const int c8 = chunk_of(hf), ch0 = c8 * 8;
      float hsum[8];                                    // fused head: partial dot products of this thread's 8 channels
#pragma unroll
      for (int d = 0; d < 8; ++d) hsum[d] = 0.f;
      mbar_wait_sleep(&sh->mma_bar[hf], mpar);
      tc_fence_after();
      if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[oi * 16 + (hf ? 9 : 5)] = clock64();
      if (!epi || ((quad * 32) >> op.log2L) >= nb) continue;   // the MMA warp; or none of this warp's 32 rows belongs to a trajectory (L = 8 ops, last CTA)
#pragma unroll 1
      for (int o = 0; o < n_out; ++o) {
        float v[CH_EC];
#pragma unroll
        for (int c = 0; c < CH_EC; ++c) v[c] = sh->vec[oi][0][ch0 + c];
        // ---- combine the tap blocks: out[l] += Y_t[l + shift] (zero outside the trajectory) ----
        // ---- combine the tap blocks: out[l] += Y_t[l + shift] (zero outside the trajectory); the op kind is uniform over the CTA ----
        const uint32_t tbase = taddr + hf * CH_HCOLS + (CL == 4 ? 0 : (slice >> 1) * op.T * 16) + (slice & 1) * 8;   // accumulator columns: [quarter][tap][16]
        if (op.kind == CH_UP) {        // ConvTranspose1d(k4, s2, p1): out[2m] = Y1[m] + Y3[m-1];  out[2m+1] = Y0[m+1] + Y2[m]
          if (o == 0) tap_combine<2>(v, tbase, {1, 3}, {0, -1}, l, L, lane);
          else tap_combine<2>(v, tbase, {0, 2}, {1, 0}, l, L, lane);
        } else if (op.T == 5) {
          tap_combine<5>(v, tbase, {0, 1, 2, 3, 4}, {-2, -1, 0, 1, 2}, l, L, lane);
        } else if (op.T == 3) {
          tap_combine<3>(v, tbase, {0, 1, 2}, {-1, 0, 1}, l, L, lane);
        } else {
          tap_combine<1>(v, tbase, {0}, {0}, l, L, lane);
        }
        if (op.gn) {
          if (L == 16) group_norm8_mish<16>(v, sh->vec[oi][1] + ch0, sh->vec[oi][2] + ch0);
          else group_norm8_mish<8>(v, sh->vec[oi][1] + ch0, sh->vec[oi][2] + ch0);
        }
#pragma unroll
        for (int c = 0; c < CH_EC; ++c) v[c] += addv[hf][c];
        if (op.res_kind == CH_RES_SMEM) {     // identity residual: the block input, still resident (hi + lo)
          const uint8_t* rb = smem + op.res_buf * CH_ACT_BYTES;
          const uint4 hh = *reinterpret_cast<const uint4*>(rb + swz(r, c8));
          const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&hh);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += __bfloat162float(hp[i]);
          if (NSPLIT == 2) {
            const uint4 ll = *reinterpret_cast<const uint4*>(rb + CH_HALF + swz(r, c8));
            const __nv_bfloat16* lp = reinterpret_cast<const __nv_bfloat16*>(&ll);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += __bfloat162float(lp[i]);
          }
        }
        // ---- output ----
        if (op.out_buf >= 0) {
          // next op's A operand, written in the swizzled K-major layout; rows without a trajectory are zero
          const int ro = op.kind == CH_UP ? sidx * 2 * L + 2 * l + o : r;
          if (op.kind != CH_UP || r < 64) {
            uint8_t* ob = smem + op.out_buf * CH_ACT_BYTES;
            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
            if (row_ok) split8(v, &hi, &lo);
            *reinterpret_cast<uint4*>(ob + swz(ro, c8)) = hi;
            if (NSPLIT == 2) *reinterpret_cast<uint4*>(ob + CH_HALF + swz(ro, c8)) = lo;
#pragma unroll
            for (int pr = 0; pr < NPEER; ++pr) {   // the same chunks into the peer CTAs' copies of the buffer
              const uint32_t ra = peer_base[pr] + (uint32_t)(op.out_buf * CH_ACT_BYTES) + swz(ro, c8);
              if (CH_STASYNC) {
                const uint32_t rb = peer_base[pr] + (smem_u32(&sh->abar[(oi + 1) & 1]) - smem_u32(smem));
                st_async_v4(ra, hi, rb);
                if (NSPLIT == 2) st_async_v4(ra + CH_HALF, lo, rb);
              } else {
                st_cluster_v4(ra, hi);
                if (NSPLIT == 2) st_cluster_v4(ra + CH_HALF, lo);
              }
            }
          }
        } else if (op.out_buf == CH_OUT_GLOBAL) {
          const bool emit = row_ok && (op.kind != CH_DOWN || (l & 1) == 0);
          if (emit) {
            const size_t orow = op.kind == CH_DOWN ? (size_t)b * (L >> 1) + (l >> 1) : (op.kind == CH_UP ? (size_t)b * 2 * L + 2 * l + o : (size_t)b * L + l);
            uint4 hi, lo;
            split8(v, &hi, &lo);
            *reinterpret_cast<uint4*>(a.out_hi + orow * 64 + ch0) = hi;
            if (NSPLIT == 2) *reinterpret_cast<uint4*>(a.out_lo + orow * 64 + ch0) = lo;
          }
        } else {
          // fused 1x1 head (final_conv.1 / act_conv.1): this thread's share of the dot products
          const int hd = a.head_dim;
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            if (d < hd) {
              const float4 w0 = *reinterpret_cast<const float4*>(sh->headw + d * 64 + ch0), w1 = *reinterpret_cast<const float4*>(sh->headw + d * 64 + ch0 + 4);
              float s = fmaf(v[0], w0.x, fmaf(v[1], w0.y, fmaf(v[2], w0.z, fmaf(v[3], w0.w, hsum[d]))));
              hsum[d] = fmaf(v[4], w1.x, fmaf(v[5], w1.y, fmaf(v[6], w1.z, fmaf(v[7], w1.w, s))));
            }
          }
        }
      }
      if (op.out_buf == CH_OUT_HEAD) {   // one partial per (column half, slice): eight of them per (row, d), whichever CTA computed them
#pragma unroll
        for (int d = 0; d < 8; ++d)
          if (d < a.head_dim) {
            headp[c8][d][r] = hsum[d];
#pragma unroll
            for (int pr = 0; pr < NPEER; ++pr) st_cluster_1f(peer_base[pr] + (uint32_t)(hb * CH_ACT_BYTES) + (uint32_t)(((c8 * 8 + d) * TC_M + r) * 4), hsum[d]);
          }
      }
    }
    mpar ^= 1;
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[oi * 16 + 8] = clock64();
    if (op.out_buf == CH_OUT_HEAD) {
      const int hd = a.head_dim;
      if (CL > 1) cluster_sync_all(); else __syncthreads();
      if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[oi * 16 + 10] = clock64();
      for (int idx = tid; idx < TC_M * hd; idx += CH_THREADS) {   // one (row, d) sum at a time; rows without a trajectory hold whatever the buffer held
        // consecutive threads take consecutive ROWS: headp[j][d][row] is row-fastest, so the loads are conflict-free (d-fastest put the hd threads of a
        // row on one bank: 7-way conflicts on all eight loads, ~1.5 k cycles per launch) and the row count is a power of two, no division
        const int d = idx >> 7, rr = idx & (TC_M - 1);
        float m = sh->headw[8 * 64 + d];
#pragma unroll
        for (int j = 0; j < CH_HEADP_SLOTS; ++j) m += headp[j][d][rr];
        sh->mo[rr * hd + d] = m;              // index rr * hd + d with rr == s * 16 + l: the [NS, H, head_dim] block of this CTA
        const int s2 = rr >> 4;
        if (a.head_out && s2 < nb && crank == 0) a.head_out[((size_t)(b0 + s2) * 16 + (rr & 15)) * hd + d] = m;
      }
      __syncthreads();
      if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[oi * 16 + 11] = clock64();
      if (a.do_sched)
#pragma unroll 1
        for (int e = tid; e < nb * HD; e += CH_THREADS) {
          // ---- fused scheduler step, one element at a time per thread: the stand-alone kernel's arithmetic and noise counters (group = 4 elements) ----
          const size_t ge = (size_t)b0 * HD + e;                // global element index
          const SchedK& k = sh->sk;
          float nz = 0.f;
          if (k.noise) nz = __ldg(k.noise + ge);
          else if (k.seed) { const float4 n4 = philox_group(*k.seed, (unsigned)(ge >> 2), k.noise_step); nz = (ge & 3) == 0 ? n4.x : ((ge & 3) == 1 ? n4.y : ((ge & 3) == 2 ? n4.z : n4.w)); }
          const float tj = k.traj ? __ldg(k.traj + ge) : 0.f;
          const float mk = k.mask ? __ldg(k.mask + ge) : 0.f;
          const int sample = e / HD, pos = e - sample * HD;
          float x0;
          const float out = sched_elem(k, sample, pos, pos % a.D, sh->mo[e], sh->xs[e], nz, tj, mk, &x0);
          if (crank == 0) a.x_out[ge] = out;
          sh->xs[e] = out;                                      // x_{t-1}: the next evaluation's input (im2col + projection)
        }
    }
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[oi * 16 + 6] = clock64();
  }

  if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[a.n_ops * 16] = clock64();
  tc_fence_before();
  if (CL > 1) cluster_sync_all();     // neither CTA leaves while the other may still store into its shared memory
  else __syncthreads();
  if (warp == CH_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int launch_chain64(const ChainArgs& a, int nsplit, cudaStream_t s) {
  if ((a.ns != 4 && a.ns != 8) || a.n_ops < 1 || a.n_ops > CH_MAXOPS || a.B <= 0 || a.H != 16 || a.D < 1 || a.D > 8 || (a.H * a.D) % 4 != 0) return B2P_ERR_INVALID_ARG;
  for (int i = 0; i < a.n_ops; ++i) {
    const ChainOp& op = a.ops[i];
    if (op.T < 1 || op.T > 5 || (op.L != 8 && op.L != 16) || op.in_buf > 2 || op.out_buf > 2 || op.ksteps < 1 || op.ksteps > 4) return B2P_ERR_INVALID_ARG;
    if (op.out_buf >= 0 && (op.out_buf == op.in_buf || (op.res_kind == CH_RES_SMEM && op.out_buf == op.res_buf))) return B2P_ERR_INVALID_ARG;
    if (op.kind == CH_UP && (op.L != 8 || op.T != 4)) return B2P_ERR_INVALID_ARG;
    if (op.out_buf == CH_OUT_HEAD && (!a.headW || a.head_dim < 1 || a.head_dim > 8 || op.L != 16)) return B2P_ERR_INVALID_ARG;
    // the head's partial sums alias activation buffer 2 (0 if the A operand is buffer 2): it must not be the residual, and the next op must rebuild its own operand
    if (op.out_buf == CH_OUT_HEAD && ((op.res_kind == CH_RES_SMEM && op.res_buf == (op.in_buf == 2 ? 0 : 2)) || (i + 1 < a.n_ops && a.ops[i + 1].in_buf != CH_IN_IM2COL))) return B2P_ERR_INVALID_ARG;
    if (op.in_buf == CH_IN_IM2COL && op.L != 16) return B2P_ERR_INVALID_ARG;
    if (i > 0 && op.in_buf >= 0 && a.ops[i - 1].out_buf != op.in_buf) return B2P_ERR_INVALID_ARG;   // the cluster forms count the previous op's stores as this op's operand
  }
  if (a.do_sched && (a.head_dim != a.D || !a.x_out || a.sk.mo_u || a.sk.clip_mode == 3)) return B2P_ERR_INVALID_ARG;
  const size_t smem = chain64_smem_bytes();
  const int groups = (a.B + a.ns - 1) / a.ns;
  // CL CTAs per trajectory group (each computes 1/CL of the channels of every op): two while all clusters are resident at once (four measured no
  // faster than two: the single epilogue pass is latency-bound by then and the barrier among four CTAs costs more); B2P_CHAIN_CL = 1 / 2 / 4 forces a variant
  static int forced = -1, max_clusters[5] = {0, 0, 0, 0, 0};
  if (forced < 0) {
    const char* e = getenv("B2P_CHAIN_CL");
    forced = e ? atoi(e) : 0;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    max_clusters[1] = sms;
    for (int c = 2; c <= 4; c += 2) {   // how many clusters of c CTAs of this footprint the device holds at once (GPC granularity)
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(c * sms); q.blockDim = dim3(CH_THREADS); q.dynamicSmemBytes = smem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = c; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int n = 0;
      cudaError_t e1 = c == 2 ? cudaFuncSetAttribute(chain64_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : cudaFuncSetAttribute(chain64_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e1 == cudaSuccess) e1 = c == 2 ? cudaOccupancyMaxActiveClusters(&n, chain64_kernel<2, 2>, &q) : cudaOccupancyMaxActiveClusters(&n, chain64_kernel<2, 4>, &q);
      if (e1 != cudaSuccess) { cudaGetLastError(); n = 0; }
      max_clusters[c] = n;
    }
    if (getenv("B2P_CHAIN_DEBUG")) fprintf(stderr, "[b2p] chain kernel: %d SMs, resident clusters of 2 CTAs: %d, of 4 CTAs: %d\n", sms, max_clusters[2], max_clusters[4]);
  }
  const int cl = forced == 1 || forced == 2 || forced == 4 ? forced : (groups <= max_clusters[2] ? 2 : 1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(groups * cl); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[3];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  int na = 1;
  if (cl > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cl; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = add_l2_window_attr(attr, na);
#define B2P_CHAIN_LAUNCH(NS_, CL_)                                                                                                  \
  do {                                                                                                                               \
    B2P_CUDA_TRY(cudaFuncSetAttribute(chain64_kernel<NS_, CL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    return (int)cudaLaunchKernelEx(&cfg, chain64_kernel<NS_, CL_>, a);                                                               \
  } while (0)
  if (nsplit == 2) { if (cl == 4) B2P_CHAIN_LAUNCH(2, 4); else if (cl == 2) B2P_CHAIN_LAUNCH(2, 2); else B2P_CHAIN_LAUNCH(2, 1); }
  if (cl == 4) B2P_CHAIN_LAUNCH(1, 4); else if (cl == 2) B2P_CHAIN_LAUNCH(1, 2); else B2P_CHAIN_LAUNCH(1, 1);
#undef B2P_CHAIN_LAUNCH
}

}  // namespace b2p
