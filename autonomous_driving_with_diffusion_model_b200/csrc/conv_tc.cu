// tcgen05 / TMEM / TMA implicit-GEMM conv with the Conv1dBlock + residual-block epilogue fused (K2 in SURVEY.md
// Appendix C).  Replaces the same reference statements as conv_ffma.cu (modeling/helpers.py:95-112,
// modeling/temporal.py:53-55, 227, 233-245) for every layer whose channel counts are multiples of 64.
//
//   Y_t[rows, Cout] = A[rows, Cin] * W_t[Cout, Cin]^T  for every tap t;   out[l] = sum_t Y_t[l + t - pad]
//   rows = (sample, position), channels-last.
//
// * A (activations) and B (weights) are bf16, K-major, staged by TMA into 128B-swizzled shared memory: A as boxes of a
//   3-D tensor map (C, L, B), the weights as boxes {64 ch, TN out-channels, T taps} of a map (Cin, Cout, taps).
// * "taps in N": all taps of a 64-channel K chunk are adjacent in shared memory, so ONE tcgen05.mma (cta_group::1,
//   M = 128, N = T*TN <= 256, K = 16) computes every Y_t into its own TN-column TMEM block.  The conv sum is a ROW shift
//   of the accumulator, done in the epilogue with warp shuffles (the L rows of a sample are adjacent lanes); zero padding
//   is "source lane outside the sample".  The activation tile is loaded once per chunk, not once per tap; no im2col.
// * precision modes: NSPLIT = 1 -> single bf16 pass;  NSPLIT = 2 -> operands stored as bf16 hi/lo pairs, the products
//   hi*hi + lo*hi + hi*lo accumulate in fp32 in TMEM ("bf16x3", fp32-class parity).  W_hi and W_lo are adjacent in N, so
//   they are TWO instructions: A_hi x [W_hi | W_lo] (the hi*lo products in their own TMEM blocks) and A_lo x W_hi — an
//   M = 128 MMA fetches its operands from shared memory at ~64 B/clk, so with narrow N the 4 KB A slice is its cost.
// * tile = 128 rows x TN output channels, TN in {64, 32, 16}: at small batch a layer has few row tiles, so narrow
//   column tiles are what spreads it over the 148 SMs.  When a GroupNorm group (Cout/8 channels) is wider than TN the
//   CTAs that share it form a thread-block cluster along N and push their per-row partial statistics into each other's
//   shared memory with st.async stores that signal the receiver's mbarrier (no cluster-wide barrier on the critical path).
// * epilogue on all 16 warps: warp w owns TMEM lane quadrant (w & 3) and column slice (w >> 2) of the tile: tap combine,
//   bias, GroupNorm(8) by warp shuffles over the L rows of a sample (+ smem / DSMEM exchange between column slices /
//   CTAs), Mish (SFU), + time embedding terms, + residual (identity or a 1x1 conv accumulated in a further TMEM block),
//   optional fused 1x1 head, bf16 hi/lo store.  Post-norm addends are prefetched before the accumulator wait.
// * programmatic dependent launch: barrier init, TMEM allocation, epilogue-vector staging and the first ring pass of
//   WEIGHT loads run before griddepcontrol.wait and overlap the previous layer; only activation loads wait for it.
// Thread 0 is the TMA producer and thread 32 the MMA issuer before they join the epilogue; warp 1 owns the TMEM allocation.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace b2p {

constexpr int TC_THREADS = 512;
constexpr int A_BYTES = TC_M * TC_K * 2;   // 16 KB

#ifdef B2P_TC_TRACE   // developer tracing (scripts/tc_trace.py): per-stage clocks of CTA (0,0) of every launch
__device__ unsigned long long tc_trace[8192 * 16];
int tc_trace_launch = 0;
#define TC_T(k)                                                                                       \
  do {                                                                                                \
    if (blockIdx.x == 0 && blockIdx.y == 0) {                                                         \
      unsigned long long t_;                                                                          \
      if ((k) < 7) t_ = clock64(); else asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));      \
      tc_trace[((a.dbg >> 8) & 8191) * 16 + (k)] = t_;                                                \
    }                                                                                                 \
  } while (0)
// slot 8: latest end over ALL CTAs of the launch, slot 9: dependency-wait return of CTA (0,0), both on the global timer
#define TC_TG(k, all)                                                                                 \
  do {                                                                                                \
    if ((all) || (blockIdx.x == 0 && blockIdx.y == 0)) {                                              \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                         \
      atomicMax(&tc_trace[((a.dbg >> 8) & 8191) * 16 + (k)], t_);                                     \
    }                                                                                                 \
  } while (0)
#else
#define TC_T(k)
#define TC_TG(k, all)
#endif

struct __align__(16) TcShared {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full;
  uint64_t gn_bar;                                 // counts the GroupNorm partials arriving from the cluster peers
  uint32_t tmem_base;
  uint32_t pad;
  float bias[64], gamma[64], beta[64], resb[64];   // per-tile epilogue vectors (first TN entries used)
};
// Shared-memory map (dynamic, 1024-aligned base):
//   [0, 224 KB)            TMA ring (TN == 64) — or ring in [0, 208 KB) and, for TN < 64, the cluster exchange buffer
//                          cx[2][4][128] (GroupNorm partials written by PEER CTAs, so it may never alias live stages) at 208 KB
//   ring start, reused after the main loop:  xchg[2][4][128] (partials between the column slices of this CTA), head[4][8][128], head weights
//   [224 KB, ...)          TcShared (barriers, TMEM base, epilogue vectors)
//   TN == 16 uses a 106 KB ring (+ cx) so that TWO CTAs fit on an SM: a 128-CTA layer then leaves room for the next layer's
//   CTAs to become resident early (programmatic dependent launch) instead of waiting for SMs to drain.
template <int TN> struct SmemPlan {
  static constexpr int ring = TN == 64 ? 224 * 1024 : (TN == 32 ? 208 * 1024 : 106 * 1024);
  static constexpr int cx_off = ring;                                      // cluster exchange buffer (TN < 64), 4 KB
  static constexpr int shared_off = TN == 64 ? ring : ring + 4 * 1024;     // TcShared
  static constexpr int total = shared_off + 1024 /*alignment slack*/ + (int)sizeof(TcShared);
  static constexpr int min_ctas = 1;   // TN == 16 launches are one wave of <= 148 CTAs with the deep ring (one CTA per SM): the full register file per CTA
};
constexpr int EPI_XCHG_BYTES = 2 * TC_M * 4 * 4;
constexpr int EPI_STAGE_OFF = 8 * 1024;            // output staging tiles (TcArgs.tma_out) in the idle operand ring, past the slice-exchange buffer: 2 planes x 128 rows x TN x 2 B
struct EpiScratch {
  float (*xchg)[4][TC_M];   // [pass][slice][row]: row fastest, so the lanes of a warp hit distinct banks
  float2 (*cx)[TC_M];       // [source CTA][row] = (mean, M2), written by the peers (st.async)
  uint64_t* gn_bar;
};

// GroupNorm(8) + Mish for the EC channels a thread holds.  A group is CG consecutive channels x the L rows (adjacent
// lanes) of a sample; TN is the CTA's column-tile width.  Every thread first computes the mean and the centred sum of
// squares (M2) of ITS part of the group (W channels x L rows, two passes over registers + xor shuffles).  When the group
// spans several column slices of the CTA (SL) and/or several CTAs of the cluster (CN), the (mean, M2) pairs of the equally
// sized parts are exchanged ONCE — through shared memory between slices, through distributed shared memory between
// CTAs — and merged with the parallel-variance formula  M2 = sum M2_p + n_p * sum (mean_p - mean)^2, which is as accurate
// as the reference's two-pass GroupNorm.  Parts are summed in a fixed order so all CTAs get bit-identical statistics.
// Called by ALL threads of the CTA (contains __syncthreads / a cluster barrier).
__device__ __forceinline__ float pow2_recip(int p) { return __uint_as_float((uint32_t)(127 - (31 - __clz(p))) << 23); }   // 1 / p for a power of two p

template <int CG, int TN>
__device__ __forceinline__ void group_norm_mish(float (&v)[TN / 4], int L, float inv_L, int row, int slice, int crank, int cbase, const EpiScratch& es,
                                                const float* gamma, const float* beta, int tr = -1) {
#ifdef B2P_TC_TRACE
#define GN_T(k) do { if (tr >= 0 && threadIdx.x == 64) tc_trace[tr * 16 + (k)] = clock64(); } while (0)
#else
#define GN_T(k)
#endif
  constexpr int EC = TN / 4;
  constexpr int W = CG < EC ? CG : EC;              // channels of one group inside this thread
  constexpr int NG = EC / W;                        // groups per thread
  constexpr int WC = CG < TN ? CG : TN;             // channels of one group inside this CTA
  constexpr int SL = WC / W;                        // slices of this CTA sharing a group
  constexpr int CN = CG / WC;                       // CTAs sharing a group
  // W, CG and L are powers of two: their reciprocals are exponent arithmetic (exact), not divisions — the divisions of the four instantiations used
  // to be hoisted to the point right after the accumulator wait, i.e. onto the critical path of every layer
  const float inv_np = inv_L * (1.0f / W);           // 1 / elements of one part (inv_L = 1 / L is computed once, before the accumulator wait)
  float mean[NG], m2[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < W; ++c) s += v[g * W + c];
    for (int o = 1; o < L; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean[g] = s * inv_np;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < W; ++c) { float d = v[g * W + c] - mean[g]; q = fmaf(d, d, q); }
    for (int o = 1; o < L; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    m2[g] = q;
  }
  GN_T(11);
  if (SL > 1) {                                      // merge the slices of this CTA (NG == 1 here)
    es.xchg[0][slice][row] = mean[0];
    es.xchg[1][slice][row] = m2[0];
    // only the four warps that own these 32 rows (same TMEM lane quadrant, the four column slices) exchange: a 128-thread named
    // barrier per quadrant instead of a CTA-wide one, so a quadrant does not wait for the slowest warp of the other three
    asm volatile("bar.sync %0, 128;" ::"r"(1 + (row >> 5)) : "memory");
    const int base = slice & ~(SL - 1);
    float ms = 0.f, qs = 0.f;
#pragma unroll
    for (int j = 0; j < SL; ++j) ms += es.xchg[0][base + j][row];
    const float mu = ms * (1.0f / SL);
#pragma unroll
    for (int j = 0; j < SL; ++j) { float d = es.xchg[0][base + j][row] - mu; qs += es.xchg[1][base + j][row] + (float)(W * L) * d * d; }
    mean[0] = mu; m2[0] = qs;
  }
  GN_T(12);
  if (CN > 1) {                                      // merge the CTAs of the cluster sub-group
    // every CTA pushes its per-row (mean, M2) into all CN CTAs with asynchronous stores that signal the receiver's
    // mbarrier; nobody waits for a cluster-wide barrier, each CTA only waits until ITS CN x 128 pairs have landed
    if ((slice & (SL - 1)) == 0) {
      const uint32_t dst = smem_u32(&es.cx[crank][row]), bar = smem_u32(es.gn_bar);
#pragma unroll
      for (int c = 0; c < CN; ++c) st_async_cluster_f32x2(dst, bar, (uint32_t)(cbase + c), mean[0], m2[0]);
    }
    mbar_wait(es.gn_bar, 0);
    float ms = 0.f, qs = 0.f;
#pragma unroll
    for (int c = 0; c < CN; ++c) ms += es.cx[c][row].x;
    const float mu = ms * (1.0f / CN);
#pragma unroll
    for (int c = 0; c < CN; ++c) { float d = es.cx[c][row].x - mu; qs += es.cx[c][row].y + (float)(WC * L) * d * d; }
    mean[0] = mu; m2[0] = qs;
  }
  GN_T(13);
  const float inv_n = inv_L * (1.0f / CG);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const float r = rsqrtf(m2[g] * inv_n + 1e-5f);
#pragma unroll
    for (int c = 0; c < W; ++c) v[g * W + c] = mish_fast((v[g * W + c] - mean[g]) * r * gamma[g * W + c] + beta[g * W + c]);
  }
}

template <int NSPLIT, int TN>
__global__ void __launch_bounds__(TC_THREADS, SmemPlan<TN>::min_ctas) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcArgs a) {
  constexpr int EC = TN / 4;                        // columns per epilogue thread
  constexpr int BT_BYTES = TN * TC_K * 2;           // one tap's weight tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  TcShared* sh = reinterpret_cast<TcShared*>(smem + (TN == 64 ? a.ring : a.ring + 4 * 1024));

  const int T = a.T;
  const int stage_bytes = NSPLIT * (A_BYTES + T * BT_BYTES);   // [A hi | A lo | W hi (T taps) | W lo (T taps)]
  const int stages = a.stages;                                   // min(ring / stage_bytes, 8), computed by the host (no division here)
  const int TB = (NSPLIT == 2 && a.concat) ? 2 * T : T;        // tap blocks in TMEM (hi*lo products separate when concatenated)
  const int need_cols = (TB + 1) * TN;
  const uint32_t tmem_cols = need_cols <= 32 ? 32u : (need_cols <= 64 ? 64u : (need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, n0 = blockIdx.y * TN;
  // cluster = (1, CL, 1) consecutive column tiles of the same row tile: they share the activation tile (TMA multicast: every
  // CTA loads 1/CL of it and broadcasts) and, in sub-groups of cluster_n, a GroupNorm group (statistics exchange)
  const int CL = a.cluster_l;
  const int lrank = blockIdx.y & (CL - 1);         // rank inside the cluster (== %cluster_ctarank); cluster sizes are powers of two
  const int crank = lrank & (a.cluster_n - 1);     // rank inside the GroupNorm sub-group
  const int cbase = lrank - crank;                 // first cluster rank of the sub-group
  const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
  // weight multicast: cluster = (CM, 1, 1) consecutive ROW tiles of the same column tile (only with CL == 1)
  const int CM = a.cluster_m;
  const bool wm = CM > 1;
  const int mrank = blockIdx.x & (CM - 1);
  const uint16_t wmask = (uint16_t)((1u << CM) - 1u);
  const int b0 = tile_m * a.samples_per_tile;

  const int chunks0 = a.C[0] / TC_K, chunks1 = a.C[1] / TC_K;
  const int main_iters = chunks0 + chunks1;
  const int rchunks0 = a.RC[0] / TC_K, rchunks1 = a.RC[1] / TC_K;
  const int res_iters = rchunks0 + rchunks1;
  const int n_local = (a.dbg & 1) ? 0 : main_iters + res_iters;

  if (threadIdx.x == 0) {
    TC_T(0);
    for (int s = 0; s < stages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], wm ? CM : ((CL > a.cluster_n) ? CL : 1)); }
    mbar_init(&sh->tmem_full, 1);
    mbar_init(&sh->gn_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (a.cluster_n > 1 && a.gn_gamma) mbar_expect_tx(&sh->gn_bar, (uint32_t)a.cluster_n * TC_M * 8u);   // one (mean, M2) pair per row from every CTA of the sub-group
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + TN) {   // stage the tile's epilogue vectors (weights: independent of the previous layer)
    const int c = threadIdx.x - 64;
    sh->bias[c] = a.bias ? __ldg(a.bias + n0 + c) : 0.f;
    sh->gamma[c] = a.gn_gamma ? __ldg(a.gn_gamma + n0 + c) : 1.f;
    sh->beta[c] = a.gn_gamma ? __ldg(a.gn_beta + n0 + c) : 0.f;
    sh->resb[c] = a.resB ? __ldg(a.resB + n0 + c) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (a.cluster_n > 1 && a.gn_gamma) cluster_sync_all();   // every peer's GroupNorm mbarrier is ready before anyone sends (in the prologue: off the dependency chain)
  const bool mcast = CL > a.cluster_n;             // activation multicast active (off by default: measured slower)
  if (mcast || wm) cluster_sync_all();             // peers' barriers are initialised before anyone multicasts / commits into them
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    if (elect_one()) {
    // =============================== TMA producer (one elected lane of warp 0: the loads compile to plain UTMALDGs) ===============================
    prefetch_tmap(&maps.a[0][0]);
    prefetch_tmap(&maps.w[0]);
    if (NSPLIT == 2) { prefetch_tmap(&maps.a[0][1]); prefetch_tmap(&maps.w[1]); }
    // The weights do not depend on the previous layer: their TMA loads for the first ring pass are issued BEFORE the
    // grid-dependency wait and overlap the tail of the preceding kernel; only the activation loads wait for it.
    const int mc_sub = mcast ? a.samples_per_tile / CL : 0, mc_bytes = mcast ? A_BYTES / CL : 0;   // multicast slices (off the hot path: once)
    const uint64_t pol_w = l2_policy_evict_last(), pol_a = l2_policy_evict_first();
    auto issue = [&](int it, int s, bool do_w, bool do_a) {   // s = it % stages, tracked by the callers (no runtime division)
      uint8_t* st = smem + s * stage_bytes;
      uint8_t* sb = st + NSPLIT * A_BYTES;
      const bool res_phase = it >= main_iters;
      const int ch = res_phase ? it - main_iters : it;
      const int nch0 = res_phase ? rchunks0 : chunks0;
      const int src = ch < nch0 ? 0 : 1;
      const int c0 = (src == 0 ? ch : ch - nch0) * TC_K;
      const int kglob = (src == 0 ? 0 : (res_phase ? a.RC[0] : a.C[0])) + c0;
      if (do_w) {
        mbar_expect_tx(&sh->full[s], res_phase ? NSPLIT * (A_BYTES + BT_BYTES) : stage_bytes);
        if (wm) {   // one box per (plane, tap), dealt round-robin to the CTAs of the cluster; each box lands in every CTA
#pragma unroll
          for (int h = 0; h < NSPLIT; ++h) {
            if (res_phase) {
              if ((h & (CM - 1)) == mrank) tma_load_2d_mc(sb + h * T * BT_BYTES, &maps.rw[h], &sh->full[s], kglob, n0, wmask);
            } else {
              for (int t = 0; t < T; ++t)
                if (((h * T + t) & (CM - 1)) == mrank) tma_load_3d_mc(sb + (h * T + t) * BT_BYTES, &maps.w[h], &sh->full[s], kglob, n0, a.tap0 + t, wmask);
            }
          }
        } else
#pragma unroll
        for (int h = 0; h < NSPLIT; ++h) {
          if (a.w_hint) {
            if (res_phase) tma_load_2d_hint(sb + h * T * BT_BYTES, &maps.rw[h], &sh->full[s], kglob, n0, pol_w);
            else tma_load_3d_hint(sb + h * T * BT_BYTES, &maps.w[h], &sh->full[s], kglob, n0, a.tap0, pol_w);
          } else if (res_phase) tma_load_2d(sb + h * T * BT_BYTES, &maps.rw[h], &sh->full[s], kglob, n0);
          else tma_load_3d(sb + h * T * BT_BYTES, &maps.w[h], &sh->full[s], kglob, n0, a.tap0);
        }
      }
      if (do_a) {
        // this CTA fetches samples [lrank*spt/CL, (lrank+1)*spt/CL) of the tile and broadcasts them to the whole cluster
#pragma unroll
        for (int h = 0; h < NSPLIT; ++h) {
          const CUtensorMap* mp = res_phase ? &maps.r[src][h] : &maps.a[src][h];
          if (mcast) tma_load_3d_mc(st + h * A_BYTES + lrank * mc_bytes, mp, &sh->full[s], c0, 0, b0 + lrank * mc_sub, cmask);
          else if (a.w_hint == 2) tma_load_3d_hint(st + h * A_BYTES, mp, &sh->full[s], c0, 0, b0, pol_a);
          else tma_load_3d(st + h * A_BYTES, mp, &sh->full[s], c0, 0, b0);
        }
      }
    };
    const int npre = n_local < stages ? n_local : stages;
    for (int j = 0; j < npre; ++j) issue(j, j, true, false);
    griddep_wait();
    TC_T(1); TC_TG(9, false);
    for (int j = 0; j < npre; ++j) issue(j, j, false, true);
    for (int j = npre, s = 0, par = 0; j < n_local; ++j) {      // s = j % stages, par = ((j / stages) & 1) ^ 1 (npre == stages here)
      mbar_wait(&sh->empty[s], par);
      issue(j, s, true, true);
      if (++s == stages) { s = 0; par ^= 1; }
    }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (one elected lane of warp 1 runs the whole loop: tc_ptx.cuh elect_one) ===============================
    // All taps of a chunk are adjacent in shared memory ([T*TN rows] x 128 B, K-major), so ONE tcgen05.mma with
    // N = T*TN (<= 256; a fifth 64-wide tap takes a second instruction) covers them.  Descriptors are built once per
    // stage and advanced by adding to the address field.
    const int nA = (T * TN <= 256) ? T : 4;                              // taps covered by the first instruction
    const uint32_t idescA = umma_idesc_n(nA * TN), idescB = umma_idesc_n(TN), idescR = umma_idesc_n((T - nA) * TN);   // idescR: the taps beyond the first instruction
    const bool concat = NSPLIT == 2 && a.concat;                         // W_hi and W_lo are adjacent in N: A_hi x [W_hi | W_lo] is one instruction
    const uint32_t idescC = umma_idesc_n(concat ? 2 * T * TN : TN);
    if (elect_one()) {
    for (int it = 0, s = 0, par = 0; it < n_local; ++it, s = (s + 1 == stages ? 0 : s + 1), par ^= (s == 0)) {   // s = it % stages, par = (it / stages) & 1
      mbar_wait(&sh->full[s], par);
      tc_fence_after();
      if (it == 0) TC_T(2);
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
      const uint32_t sb = sa + NSPLIT * A_BYTES;
      const bool res_phase = it >= main_iters;
      const bool first = res_phase ? (it == main_iters) : (it == 0);
      const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + A_BYTES);
      const uint64_t b_hi = umma_desc(sb), b_lo = umma_desc(sb + T * BT_BYTES);
      if (!res_phase) {
#pragma unroll
        for (int k = 0; k < TC_K / TC_UMMA_K; ++k) {
          const uint64_t ko = (uint64_t)(k * TC_UMMA_K * 2 / 16);       // +32 B per K step, in 16-byte units
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          if (concat) {
            // an M = 128 MMA fetches its operands from shared memory at ~64 B/clk: with N = 48 the 4 KB A slice dominates, so
            // two instructions that read A_hi once and A_lo once beat three that read A_hi twice
            umma(tmem_base, a_hi + ko, b_hi + ko, idescC, acc);          // blocks [0,T): hi*hi, blocks [T,2T): hi*lo
            umma(tmem_base, a_lo + ko, b_hi + ko, idescA, 1u);           // blocks [0,T) += lo*hi
            continue;
          }
          umma(tmem_base, a_hi + ko, b_hi + ko, idescA, acc);
          if (NSPLIT == 2) {
            umma(tmem_base, a_lo + ko, b_hi + ko, idescA, 1u);
            umma(tmem_base, a_hi + ko, b_lo + ko, idescA, 1u);
          }
          if (nA < T) {
            const uint64_t t4 = (uint64_t)(nA * BT_BYTES / 16);
            umma(tmem_base + nA * TN, a_hi + ko, b_hi + t4 + ko, idescR, acc);
            if (NSPLIT == 2) {
              umma(tmem_base + nA * TN, a_lo + ko, b_hi + t4 + ko, idescR, 1u);
              umma(tmem_base + nA * TN, a_hi + ko, b_lo + t4 + ko, idescR, 1u);
            }
          }
        }
      } else {
        const uint32_t d = tmem_base + TB * TN;
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
      if (wm) umma_commit_mc(&sh->empty[s], wmask); else if (CL > a.cluster_n) umma_commit_mc(&sh->empty[s], cmask); else umma_commit(&sh->empty[s]);   // the stage is free in a CTA once ALL cluster consumers released it
    }
    umma_commit(&sh->tmem_full);
    TC_T(3);
    }
  }
  __syncwarp();
  griddep_wait();                 // everything below may read the previous kernels' outputs (residuals)
  griddep_launch_dependents();    // the next layer may start its prologue / weight prefetch on the idle SMs

  // =============================== epilogue (all 16 warps) ===============================
  {
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int slice = warp >> 2;                    // EC-column slice of the tile
    const int col0 = slice * EC;
    const int r = quad * 32 + lane;                 // tile row == TMEM lane
    const int L = a.Lrows;
    const long grow = (long)tile_m * TC_M + r;
    const bool row_ok = grow < a.nrows;
    const int b = (int)(grow >> a.log2L), l = (int)(grow & (L - 1));
    EpiScratch es;
    es.xchg = reinterpret_cast<float (*)[4][TC_M]>(smem);
    es.cx = reinterpret_cast<float2 (*)[TC_M]>(smem + a.ring);
    es.gn_bar = &sh->gn_bar;
    float (*headp)[8][TC_M] = reinterpret_cast<float (*)[8][TC_M]>(smem + EPI_XCHG_BYTES);   // [slice][d][row]: row fastest, conflict-free
    float* headw = reinterpret_cast<float*>(smem + EPI_XCHG_BYTES + TC_M * 4 * 8 * 4);   // [8][64] head weights + [8] bias, after headp
    const int gcol = n0 + col0;
    const int n_out = (a.dbg & 2) ? 0 : a.n_out;
    const bool has_res = res_iters > 0;
    // everything that is added AFTER GroupNorm/Mish (time embedding terms, identity residual, residual-conv bias) is
    // fetched from global memory now, while the tensor core is still busy
    float addv[EC];
#pragma unroll
    for (int c = 0; c < EC; ++c) addv[c] = has_res ? sh->resb[col0 + c] : 0.f;
    if (row_ok && !(a.dbg & 8)) {   // dbg bit 3: skip the addend loads (timing experiment)
      if (a.temb) {
        const float* tp = a.temb + (size_t)b * a.temb_stride + gcol;
#pragma unroll
        for (int c = 0; c < EC; c += 4) { float4 t4 = __ldg(reinterpret_cast<const float4*>(tp + c)); addv[c] += t4.x; addv[c + 1] += t4.y; addv[c + 2] += t4.z; addv[c + 3] += t4.w; }
      }
      if (a.temb2) {
#pragma unroll
        for (int c = 0; c < EC; c += 4) { float4 t4 = __ldg(reinterpret_cast<const float4*>(a.temb2 + gcol + c)); addv[c] += t4.x; addv[c + 1] += t4.y; addv[c + 2] += t4.z; addv[c + 3] += t4.w; }
      }
      if (a.res_f32) {
        const float* q = a.res_f32 + (size_t)grow * a.Cout + gcol;
#pragma unroll
        for (int c = 0; c < EC; c += 4) { float4 t4 = __ldg(reinterpret_cast<const float4*>(q + c)); addv[c] += t4.x; addv[c + 1] += t4.y; addv[c + 2] += t4.z; addv[c + 3] += t4.w; }
      }
      if (a.res_hi) {
        const __nv_bfloat16* qh = a.res_hi + (size_t)grow * a.Cout + gcol;
        const __nv_bfloat16* ql = a.res_lo ? a.res_lo + (size_t)grow * a.Cout + gcol : nullptr;
#pragma unroll
        for (int c = 0; c < EC; c += 4) {
          uint2 hh = __ldg(reinterpret_cast<const uint2*>(qh + c));
          const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&hh);
#pragma unroll
          for (int i = 0; i < 4; ++i) addv[c + i] += __bfloat162float(hp[i]);
          if (ql) {
            uint2 ll = __ldg(reinterpret_cast<const uint2*>(ql + c));
            const __nv_bfloat16* lp = reinterpret_cast<const __nv_bfloat16*>(&ll);
#pragma unroll
            for (int i = 0; i < 4; ++i) addv[c + i] += __bfloat162float(lp[i]);
          }
        }
      }
    }
    // fused head: this thread's element of the 1x1 head weights / bias, in flight during the accumulator wait
    const float hw_reg = (a.headW && (int)threadIdx.x < 64 * a.head_dim) ? __ldg(a.headW + threadIdx.x) : 0.f;
    const float hb_reg = (a.headW && (int)threadIdx.x < a.head_dim) ? __ldg(a.headB + threadIdx.x) : 0.f;
    const float inv_L = pow2_recip(L);
    mbar_wait_sleep(&sh->tmem_full, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TC_T(4);
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + col0;

    for (int o = 0; o < n_out; ++o) {
      float v[EC];
#pragma unroll
      for (int c = 0; c < EC; ++c) v[c] = sh->bias[col0 + c];
      // ---- combine the tap blocks: out[l] += Y_t[l + shift] (zero outside the sample) ----
      if (n_local > 0) {
        if (EC == 4) {
          // narrowest tiles: EVERY tap block (and its hi*lo twin when the products are concatenated) is loaded before ONE wait — a
          // TMEM load round trip costs 700-900 cycles with 16 warps reading at once (profiles/r02_tc_stage_trace_b256.txt) — and the
          // twins are added before the row shift (the shift is linear), which halves the shuffles.  Straight-line code: taps beyond
          // nt contribute zero, so no shuffle sits behind a data-dependent branch.
          const bool twin = NSPLIT == 2 && a.concat;
          const int nt = a.nt[o];
          float y[5][EC], y2[5][EC];
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            if (i < nt) {
              tmem_ld<EC, false>(taddr + a.tap_blk[o][i] * TN, y[i]);
              if (twin) tmem_ld<EC, false>(taddr + (T + a.tap_blk[o][i]) * TN, y2[i]);
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const bool on = i < nt;
            const int d = on ? a.tap_shift[o][i] : 0;
            const bool valid = on && (l + d >= 0) && (l + d < L);
            const int src = (lane + d) & 31;
#pragma unroll
            for (int c = 0; c < EC; ++c) {
              const float t = on ? (twin ? y[i][c] + y2[i][c] : y[i][c]) : 0.f;
              const float g = __shfl_sync(0xffffffffu, t, src);
              v[c] += valid ? g : 0.f;
            }
          }
        } else if (EC <= 8) {
          // narrow tiles: issue the TMEM loads of all taps back to back and wait once
          float y[5][EC];
#pragma unroll
          for (int i = 0; i < 5; ++i)
            if (i < a.nt[o]) tmem_ld<EC, false>(taddr + a.tap_blk[o][i] * TN, y[i]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            if (i < a.nt[o]) {
              const int d = a.tap_shift[o][i];
              const bool valid = (l + d >= 0) && (l + d < L);
              const int src = (lane + d) & 31;
#pragma unroll
              for (int c = 0; c < EC; ++c) { float g = __shfl_sync(0xffffffffu, y[i][c], src); v[c] += valid ? g : 0.f; }
            }
          }
          if (NSPLIT == 2 && a.concat) {   // second round: the hi*lo products (shifts are linear, so they are added the same way)
#pragma unroll
            for (int i = 0; i < 5; ++i)
              if (i < a.nt[o]) tmem_ld<EC, false>(taddr + (T + a.tap_blk[o][i]) * TN, y[i]);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              if (i < a.nt[o]) {
                const int d = a.tap_shift[o][i];
                const bool valid = (l + d >= 0) && (l + d < L);
                const int src = (lane + d) & 31;
#pragma unroll
                for (int c = 0; c < EC; ++c) { float g = __shfl_sync(0xffffffffu, y[i][c], src); v[c] += valid ? g : 0.f; }
              }
            }
          }
        } else {
          for (int i = 0; i < a.nt[o]; ++i) {
            const int t = a.tap_blk[o][i], d = a.tap_shift[o][i];
            const bool valid = (l + d >= 0) && (l + d < L);
            const int src = (lane + d) & 31;
            float y[EC];
            tmem_ld<EC>(taddr + t * TN, y);
            if (d == 0) {
#pragma unroll
              for (int c = 0; c < EC; ++c) v[c] += y[c];
            } else {
#pragma unroll
              for (int c = 0; c < EC; ++c) { float g = __shfl_sync(0xffffffffu, y[c], src); v[c] += valid ? g : 0.f; }
            }
          }
        }
      }
#ifdef B2P_TC_TRACE
      const int gtr = (blockIdx.x == 0 && blockIdx.y == 0) ? (int)((a.dbg >> 8) & 8191) : -1;
      if (gtr >= 0 && threadIdx.x == 64) tc_trace[gtr * 16 + 10] = clock64();
#else
      const int gtr = -1;
#endif
      if (a.gn_gamma) {
        switch (a.cg) {
          case 8: group_norm_mish<8, TN>(v, L, inv_L, r, slice, crank, cbase, es, sh->gamma + col0, sh->beta + col0, gtr); break;
          case 16: group_norm_mish<16, TN>(v, L, inv_L, r, slice, crank, cbase, es, sh->gamma + col0, sh->beta + col0, gtr); break;
          case 32: group_norm_mish<32, TN>(v, L, inv_L, r, slice, crank, cbase, es, sh->gamma + col0, sh->beta + col0, gtr); break;
          default: group_norm_mish<64, TN>(v, L, inv_L, r, slice, crank, cbase, es, sh->gamma + col0, sh->beta + col0, gtr); break;
        }
      }
      const bool ok = row_ok && (l & (a.out_ldiv - 1)) == 0;   // out_ldiv is 1 or 2 (stride of the conv): masks and shifts, not divisions
      if (threadIdx.x == 64 && o == 0) TC_T(5);
#pragma unroll
      for (int c = 0; c < EC; ++c) v[c] += addv[c];
      if (has_res && n_local > 0) {   // residual 1x1 conv accumulated in the TMEM block after the tap blocks
        float rv[EC];
        tmem_ld<EC>(taddr + TB * TN, rv);
#pragma unroll
        for (int c = 0; c < EC; ++c) v[c] += rv[c];

      }
      if (a.tma_out) {
        // the tile leaves through shared memory: every thread drops its EC channels of its row into a staging tile per plane (the layout a TMA
        // box {TN channels, 128 rows} has with the 32 / 64 / 128-byte swizzle of a TN x 2-byte row), thread 0 issues one TMA store per plane after
        // the CTA-wide barrier below.  Rows past nrows are clipped by the store.  (The operand ring is idle: every MMA has retired.)
        uint8_t* st_hi = smem + EPI_STAGE_OFF;
        uint8_t* st_lo = st_hi + TC_M * TN * 2;
#pragma unroll
        for (int c = 0; c < EC; c += 4) {
          uint2 hh, ll;
          __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(&hh);
          __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(&ll);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x = v[c + i];
            hp[i] = __float2bfloat16_rn(x);
            lp[i] = __float2bfloat16_rn(x - __bfloat162float(hp[i]));
          }
          const int byte = (col0 + c) * 2;                          // byte offset inside the TN x 2-byte row
          const int chunk = byte >> 4;                                // 16-byte chunk of the row
          const int swz = TN == 64 ? (r & 7) : (TN == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));
          const uint32_t off = (uint32_t)(r * (TN * 2) + ((chunk ^ swz) << 4) + (byte & 15));
          *reinterpret_cast<uint2*>(st_hi + off) = hh;
          if (NSPLIT == 2) *reinterpret_cast<uint2*>(st_lo + off) = ll;
        }
      } else if (ok && !(a.dbg & 4)) {   // dbg bit 2: skip the global stores (timing experiment)
        const size_t orow = (size_t)b * a.out_L + (size_t)(l >> (a.out_ldiv >> 1)) * a.out_lmul + o;
        if (a.out_hi) {
          __nv_bfloat16* oh = a.out_hi + orow * a.Cout + gcol;
          __nv_bfloat16* ol = a.out_lo ? a.out_lo + orow * a.Cout + gcol : nullptr;
#pragma unroll
          for (int c = 0; c < EC; c += 4) {
            uint2 hh, ll;
            __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(&hh);
            __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(&ll);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float x = v[c + i];
              hp[i] = __float2bfloat16_rn(x);
              lp[i] = __float2bfloat16_rn(x - __bfloat162float(hp[i]));
            }
            *reinterpret_cast<uint2*>(oh + c) = hh;
            if (ol) *reinterpret_cast<uint2*>(ol + c) = ll;
          }
        }
      }
      if (a.aux_out && o == 0 && n_local > 0) {   // the extra tap block: a 1x1 conv of the same rows (residual projection), fp32 rows
        float rv[EC];
        tmem_ld<EC>(taddr + a.aux_blk * TN, rv);
        if (NSPLIT == 2 && a.concat) {
          float r2[EC];
          tmem_ld<EC>(taddr + (T + a.aux_blk) * TN, r2);
#pragma unroll
          for (int c = 0; c < EC; ++c) rv[c] += r2[c];
        }
        if (row_ok) {
          float* q = a.aux_out + (size_t)grow * a.Cout + gcol;
#pragma unroll
          for (int c = 0; c < EC; c += 4)
            *reinterpret_cast<float4*>(q + c) = make_float4(rv[c] + __ldg(a.aux_bias + gcol + c), rv[c + 1] + __ldg(a.aux_bias + gcol + c + 1),
                                                            rv[c + 2] + __ldg(a.aux_bias + gcol + c + 2), rv[c + 3] + __ldg(a.aux_bias + gcol + c + 3));
        }
      }
      if (a.headW) {   // fused 1x1 head (TN == Cout == 64): partial dot products per column slice, summed by slice 0
        // head weights as [d][64] (+ bias) in the idle operand ring (a TN == 64 launch has no peers writing into it); the
        // values were fetched before the accumulator wait, so no global latency is exposed here
        const int hd = a.head_dim;
        if (threadIdx.x < 64 * hd) { const int c = threadIdx.x / hd, d = threadIdx.x - c * hd; headw[d * 64 + c] = hw_reg; }
        if (threadIdx.x < 8) headw[8 * 64 + threadIdx.x] = hb_reg;
        __syncthreads();
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          if (d < hd) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < EC; c += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(headw + d * 64 + col0 + c);
              s = fmaf(v[c], w4.x, fmaf(v[c + 1], w4.y, fmaf(v[c + 2], w4.z, fmaf(v[c + 3], w4.w, s))));
            }
            headp[slice][d][r] = s;
          }
        }
        __syncthreads();
        if (slice == 0 && ok) {
#pragma unroll
          for (int d = 0; d < 8; ++d)
            if (d < hd) a.head_out[(size_t)grow * hd + d] = headw[8 * 64 + d] + headp[0][d][r] + headp[1][d][r] + headp[2][d][r] + headp[3][d][r];
        }
      }
    }
    tc_fence_before();
    if (a.tma_out) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows -> visible to the TMA store
    if (threadIdx.x == 64) { TC_T(6); TC_T(7); }
    if (threadIdx.x == 96) TC_TG(8, true);
  }
  if (mcast || wm) cluster_sync_all();             // no CTA may exit while peers can still signal its barriers / write its smem
  __syncthreads();
  if (a.tma_out && threadIdx.x == 0 && !(a.dbg & 4)) {
    tma_store_2d(smem + EPI_STAGE_OFF, &maps.o[0], n0, tile_m * TC_M);
    if (NSPLIT == 2) tma_store_2d(smem + EPI_STAGE_OFF + TC_M * TN * 2, &maps.o[1], n0, tile_m * TC_M);
    tma_store_commit_and_wait_read();            // the CTA may not exit while the store still reads its shared memory
  }
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
int tc_make_weight_map(CUtensorMap* m, const void* base, int taps, int Cout, int K, int box_taps, int box_n) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return B2P_ERR_NO_DEVICE;
  CUresult r;
  if (box_taps == 0) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_K, (cuuint32_t)box_n};
    cuuint32_t estr[2] = {1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Cout, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)Cout * K * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_K, (cuuint32_t)box_n, (cuuint32_t)box_taps};
    cuuint32_t estr[3] = {1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? B2P_OK : B2P_ERR_INVALID_ARG;
}

// output rows [nrows][Cout] bf16 -> box {tile_n channels, 128 rows}, swizzle = the row width of the box (32 / 64 / 128 bytes)
int tc_make_out_map(CUtensorMap* m, const void* base, int nrows, int Cout, int tile_n) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return B2P_ERR_NO_DEVICE;
  cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)nrows};
  cuuint64_t strides[1] = {(cuuint64_t)Cout * 2};
  cuuint32_t box[2] = {(cuuint32_t)tile_n, (cuuint32_t)TC_M}, estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = tile_n == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (tile_n == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2P_OK : B2P_ERR_INVALID_ARG;
}

template <int NSPLIT, int TN>
static int launch_t(const TcMaps& maps, const TcArgs& a_in, dim3 grid, cudaStream_t s) {
  TcArgs a = a_in;
  // B2P_TC_BIGRING (default 1): a narrow-tile layer that fits the chip with one CTA per SM gets the deep (208 KB, 4-stage) ring.
  // Round 1 measured no gain (the main loop was paced by the MMA issue loop); with the warp-uniform issue loop the TMA latency of
  // the 2-stage ring shows, and the deep ring is worth 1.5 % per iteration at B = 256 (0: 106 KB ring, two CTAs per SM).
  static int bigring = -1;
  if (bigring < 0) { const char* e = getenv("B2P_TC_BIGRING"); bigring = e ? atoi(e) : 1; }
  static int concat = -1;
  if (concat < 0) { const char* e = getenv("B2P_TC_CONCAT"); concat = e ? atoi(e) : 1; }
  // worth it when the A_hi re-reads it saves (64 cycles per K step, 4 K steps per 64-channel chunk) outweigh the extra TMEM
  // reads of the epilogue (T*TN more columns of 128 rows at 64 B/clk = 8*T*TN cycles): not for the 1-chunk layers
  const int chunks = (a.C[0] + a.C[1]) / TC_K;
  a.concat = (NSPLIT == 2 && concat && TN <= 32 && 2 * a.T * TN <= 256 && (concat > 1 || chunks * 256 > 10 * a.T * TN)) ? 1 : 0;
  a.ring = SmemPlan<TN>::ring;
  if (TN == 16 && bigring && (int)(grid.x * grid.y) <= 148) a.ring = SmemPlan<32>::ring;
  if (TN == 16 && 2 * NSPLIT * (A_BYTES + a.T * TN * TC_K * 2) > a.ring) a.ring = SmemPlan<32>::ring;   // six tap blocks: two stages need the deep ring
  {
    const int stage_bytes = NSPLIT * (A_BYTES + a.T * TN * TC_K * 2);
    a.stages = a.ring / stage_bytes > 8 ? 8 : a.ring / stage_bytes;
  }
  const int smem = (TN == 64 ? a.ring : a.ring + 4 * 1024) + 1024 + (int)sizeof(TcShared);
  B2P_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<NSPLIT, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemPlan<TN == 16 ? 32 : TN>::total));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[3];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.cluster_m; attr[0].val.clusterDim.y = a.cluster_l; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("B2P_TC_PDL"); pdl = e ? atoi(e) : 1; }
  int na = 1;
  if (pdl) na = 2; else attr[1] = attr[0];
  na = add_l2_window_attr(attr, na);
  cfg.attrs = attr; cfg.numAttrs = na;
  return (int)cudaLaunchKernelEx(&cfg, conv_tc_kernel<NSPLIT, TN>, maps, a);
}

// Column-tile width for a layer: the narrowest of {64, 32, 16} that still fits the launch in one wave of the 148 SMs
// (at small batch a layer has few 128-row tiles, so narrow tiles are what spreads it over the chip).
int tc_pick_tile_n(int nrows, int Cout, bool has_head) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("B2P_TC_TN"); forced = e ? atoi(e) : 0; }
  if (has_head) return 64;
  if (forced == 16 || forced == 32 || forced == 64) return forced;
  {  // developer experiment: B2P_TC_TN_MAP="<Cout>:<TN>,..." overrides the width for layers with that many output channels
    static int map_c[8], map_t[8], nmap = -1;
    if (nmap < 0) {
      nmap = 0;
      const char* e = getenv("B2P_TC_TN_MAP");
      while (e && *e && nmap < 8) {
        int c = 0, t = 0, used = 0;
        if (sscanf(e, "%d:%d%n", &c, &t, &used) == 2) { map_c[nmap] = c; map_t[nmap] = t; ++nmap; e += used; if (*e == ',') ++e; } else break;
      }
    }
    for (int i = 0; i < nmap; ++i) if (map_c[i] == Cout && (map_t[i] == 16 || map_t[i] == 32 || map_t[i] == 64)) return map_t[i];
  }
  const int mt = (nrows + TC_M - 1) / TC_M;
  int tn = 64;
  while (tn > 16 && mt * (Cout / (tn / 2)) <= 148) tn /= 2;
  return tn;
}

// Decide the tiling of one layer launch: column-tile width, GroupNorm sub-group size and multicast cluster size.
// Must be called after the layer fields (nrows, Cout, gn_gamma/cg, headW, samples_per_tile) are set and BEFORE the
// activation tensor maps are encoded (their box covers samples_per_tile / cluster_l samples).
int tc_configure(TcArgs& a) {
  a.tile_n = tc_pick_tile_n(a.nrows, a.Cout, a.headW != nullptr);
  if (a.T == 6 && a.tile_n == 64) a.tile_n = 32;   // six tap blocks of 64 columns would leave a single ring stage
  const int TN = a.tile_n;
  if (a.Cout % TN) return B2P_ERR_INVALID_ARG;
  a.cluster_n = 1;
  if (a.gn_gamma) {
    if (a.cg != 8 && a.cg != 16 && a.cg != 32 && a.cg != 64) return B2P_ERR_INVALID_ARG;
    if (a.cg > TN) a.cluster_n = a.cg / TN;          // CTAs sharing a GroupNorm group exchange statistics over DSMEM
  }
  // multicast cluster: as many column tiles as share the row tile, <= 8, a multiple of the GroupNorm sub-group, and no
  // more than the samples of a row tile (every CTA loads whole samples)
  static int mc = -1;
  if (mc < 0) { const char* e = getenv("B2P_TC_MCAST"); mc = e ? atoi(e) : 1; }   // opt-in: the lock-step it imposes costs more than the L2 reads it saves
  int cl = a.cluster_n;
  const int ntiles = a.Cout / TN;
  while (cl * 2 <= mc && cl * 2 <= 8 && ntiles % (cl * 2) == 0 && a.samples_per_tile % (cl * 2) == 0) cl *= 2;
  a.cluster_l = cl;
  // weight multicast along M (B2P_TC_WMCAST = cluster size, default below): only for launches of several waves whose column tiles need no
  // cluster of their own; the caller encodes per-tap weight boxes when cluster_m > 1
  static int wmc = -1;
  if (wmc < 0) { const char* e = getenv("B2P_TC_WMCAST"); wmc = e ? atoi(e) : 1; }
  a.cluster_m = 1;
  const int mt = (a.nrows + TC_M - 1) / TC_M;
  if (cl == 1 && wmc > 1 && mt * ntiles > 148) {
    int cm = 1;
    while (cm * 2 <= wmc && cm * 2 <= 8 && mt % (cm * 2) == 0) cm *= 2;
    a.cluster_m = cm;
  }
  return (ntiles % a.cluster_l) ? B2P_ERR_INVALID_ARG : B2P_OK;
}

// Argument errors are reported through the handle (b2p_last_error), never on stderr: the message is kept per thread until the
// caller (api.cu) attaches it to its handle.
static thread_local char tc_err[256] = "";
const char* tc_last_error() { return tc_err; }
static int tc_fail(int line, const TcArgs& a, const char* what) {
  snprintf(tc_err, sizeof(tc_err), "%s (conv_tc.cu:%d: T=%d TN=%d Cout=%d cg=%d L=%d n_out=%d cluster=%d/%d)", what, line, a.T, a.tile_n, a.Cout, a.cg,
           a.Lrows, a.n_out, a.cluster_n, a.cluster_l);
  return B2P_ERR_INVALID_ARG;
}

int launch_conv_tc(const TcMaps& maps, const TcArgs& a_in, int nsplit, cudaStream_t s) {
  TcArgs a = a_in;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("B2P_TC_DBG"); dbg = e ? atoi(e) : 0; } a.dbg = dbg; }
  { static int wh = -1; if (wh < 0) { const char* e = getenv("B2P_TC_WHINT"); wh = e ? atoi(e) : 1; } a.w_hint = wh; }
#ifdef B2P_TC_TRACE
  a.dbg |= (tc_trace_launch++ & 8191) << 8;
#endif
  const int TN = a.tile_n;
  if (TN != 64 && TN != 32 && TN != 16) return tc_fail(__LINE__, a, "invalid layer shape");
  if (a.T < 1 || a.T > 6 || a.n_out < 1 || a.n_out > 2 || (a.out_ldiv != 1 && a.out_ldiv != 2)) return tc_fail(__LINE__, a, "invalid layer shape");
  if (nsplit * (A_BYTES + a.T * TN * TC_K * 2) * 2 > (TN == 64 ? SmemPlan<64>::ring : SmemPlan<32>::ring)) return tc_fail(__LINE__, a, "invalid layer shape");
  if (a.Cout % TN || a.C[0] % TC_K || a.C[1] % TC_K || a.RC[0] % TC_K || a.RC[1] % TC_K || a.nrows <= 0) return tc_fail(__LINE__, a, "invalid layer shape");
  if (a.Lrows > 32 || (a.Lrows & (a.Lrows - 1)) || TC_M % a.Lrows) return tc_fail(__LINE__, a, "invalid layer shape");
  if (a.headW && (a.Cout != 64 || TN != 64 || a.head_dim > 8)) return tc_fail(__LINE__, a, "invalid layer shape");
  if (a.n_out == 2 && a.gn_gamma) return tc_fail(__LINE__, a, "GroupNorm with two outputs per row is not supported (single-use exchange barrier)");
  if (a.n_out == 2 && (a.RC[0] || a.RC[1])) return tc_fail(__LINE__, a, "invalid layer shape");
  if (a.tma_out && (a.n_out != 1 || a.out_ldiv != 1 || a.headW || !a.out_hi || a.out_L != a.Lrows || a.out_lmul != 1)) return tc_fail(__LINE__, a, "TMA output needs output row == GEMM row");
  if (a.cluster_n < 1 || a.cluster_l < a.cluster_n || a.cluster_l % a.cluster_n || (a.cluster_n & (a.cluster_n - 1)) || (a.cluster_l & (a.cluster_l - 1))) return tc_fail(__LINE__, a, "tc_configure() was not applied");
  if ((a.Cout / TN) % a.cluster_l) return tc_fail(__LINE__, a, "invalid layer shape");
  if (a.cluster_m < 1 || (a.cluster_m & (a.cluster_m - 1)) || a.cluster_m > 8 || (a.cluster_m > 1 && (a.cluster_l != 1 || ((a.nrows + TC_M - 1) / TC_M) % a.cluster_m)))
    return tc_fail(__LINE__, a, "tc_configure() was not applied");
  dim3 grid((a.nrows + TC_M - 1) / TC_M, a.Cout / TN, 1);
  if (nsplit == 2) {
    if (TN == 64) return launch_t<2, 64>(maps, a, grid, s);
    if (TN == 32) return launch_t<2, 32>(maps, a, grid, s);
    return launch_t<2, 16>(maps, a, grid, s);
  }
  if (TN == 64) return launch_t<1, 64>(maps, a, grid, s);
  if (TN == 32) return launch_t<1, 32>(maps, a, grid, s);
  return launch_t<1, 16>(maps, a, grid, s);
}

}  // namespace b2p

#ifdef B2P_TC_TRACE
extern "C" __attribute__((visibility("default"))) int b2p_debug_tc_trace(unsigned long long* out, int* next_launch) {
  *next_launch = b2p::tc_trace_launch;
  return (int)cudaMemcpyFromSymbol(out, b2p::tc_trace, sizeof(unsigned long long) * 8192 * 16);
}
#endif
