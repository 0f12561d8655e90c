// Small-batch fused conv layer on CUDA cores, exact fp32 (K3 in SURVEY.md Appendix C: weight-streaming GEMV).
// Same contract as conv_ffma.cu (ConvArgs; Conv1dBlock / residual block epilogue of modeling/helpers.py:95-112 and
// modeling/temporal.py:53-55), specialised for single-trajectory plan latency (the closed-loop use of the reference,
// carla agent -> generate_traj with B = 1).
//
// With L_out <= 16 rows per sample the layer is a GEMV and the work is streaming the layer's weights.  Grid =
// (C_out / NC, samples): a CTA owns NC in {1,2,4,8} output channels of one sample, NC chosen so that every layer runs on
// ~64 CTAs per sample; its 16 warps are groups of 1-2 channels x slices of K = taps x C_in (two channels per warp, sharing
// every activation read, while the launch has at most one CTA per SM).  The CTA's weight slice
// ([NC][K] fp32, from a K-major copy of the layer's weights) is fetched with cp.async BEFORE the programmatic-dependency
// wait, and the kernel releases its dependents at its very first instruction, so weight streaming runs several layers
// ahead of the dependency chain; after the wait only the (tiny) input activation is read.  A GroupNorm group (C_out/8
// channels) spans cg/NC CTAs: they form a thread-block cluster and push their (mean, M2) into each other's shared memory
// with st.async stores that signal the receiver's mbarrier; the parts are merged with the parallel-variance formula.  All
// sums are combined in a fixed order: results are run-to-run identical.
#include "common.cuh"
#include <stdlib.h>

namespace b2p {

constexpr int GV_MAXNC = 8;     // max output channels per CTA
constexpr int GV_NW = 16;       // warps per CTA = channels x K slices
constexpr int GV_NT = 32 * GV_NW;
constexpr int GV_MAXL = 16;     // max output positions per sample

#ifdef B2P_GV_TRACE   // developer tracing (scripts/gemv_bench.cu): per-stage clocks of CTA (0,0) of every launch
__device__ unsigned long long gv_trace[8192 * 8];
int gv_trace_launch = 0;
#define GV_T(k)                                                                                      \
  do {                                                                                               \
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {                                    \
      unsigned long long t_;                                                                         \
      if ((k) < 6) t_ = clock64(); else asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));     \
      gv_trace[trace_id * 8 + (k)] = t_;                                                             \
    }                                                                                                \
  } while (0)
#else
#define GV_T(k)
#endif

__device__ __forceinline__ void gv_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void gv_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 8-byte asynchronous store into CTA `cta` of the cluster that also signals the bytes on THAT CTA's mbarrier: the
// receiver needs no cluster-wide barrier, only a wait on its own mbarrier
__device__ __forceinline__ void gv_st_async(const void* local_dst, const void* local_bar, uint32_t cta, float a, float b) {
  uint32_t ld = (uint32_t)__cvta_generic_to_shared(local_dst), lb = (uint32_t)__cvta_generic_to_shared(local_bar), rd, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rd) : "r"(ld), "r"(cta));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"(cta));
  const unsigned long long v = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(rd), "l"(v), "r"(rb) : "memory");
}
__device__ __forceinline__ void gv_bar_init_expect(unsigned long long* bar, uint32_t bytes) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gv_bar_wait(unsigned long long* bar) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "GVW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra GVD_%=;\n\t"
      "bra GVW_%=;\n\t"
      "GVD_%=:\n\t}" ::"r"(a) : "memory");
}
__device__ __forceinline__ void gv_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void gv_cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct GvStatic {
  float P[GV_NW * GV_MAXL];      // per-K-slice partial conv outputs: [slice][row][channel]: (16 CW / NC) slices x L x NC <= 256 floats
  float RP[GV_NW * GV_MAXL];     // same for the residual 1x1 conv
  float O[GV_MAXL * GV_MAXNC];   // conv outputs (+bias): [row][channel]
  int src[5][GV_MAXL];           // offset (floats) into X of the input row feeding (tap, output row); zero row when padding
  int idsrc[1][GV_MAXL];         // same for the residual 1x1 conv (row r feeds row r)
  float mean, m2;
  float2 cx[8];                  // cluster exchange: (mean, M2) of every CTA of the cluster, written by the peers
  unsigned long long bar;        // mbarrier: counts the bytes of cx that have arrived
};

// x / d for a runtime d that is almost always a power of two
struct GvDiv {
  int d, sh;
  __device__ __forceinline__ int div(int x) const { return sh >= 0 ? x >> sh : x / d; }
};
__device__ __forceinline__ GvDiv gv_div(int d) {
  GvDiv r;
  r.d = d; r.sh = (d & (d - 1)) == 0 ? 31 - __clz(d) : -1;
  return r;
}

// global -> shared copy of n floats: 16-byte cp.async when both sides allow it, scalar otherwise
__device__ __forceinline__ void gv_copy(float* dst, const float* __restrict__ src, int n, bool vec, int tid) {
  if (vec) for (int i = tid * 4; i < n; i += GV_NT * 4) gv_cp16(dst + i, src + i);
  else for (int i = tid; i < n; i += GV_NT) dst[i] = __ldg(src + i);
}

// rows [nrow][C0 | C1] gathered from two channels-last sources into shared memory
__device__ __forceinline__ void gv_load_rows(float* dst, const float* __restrict__ g0, int C0, const float* __restrict__ g1, int C1,
                                             int nrow, int tid) {
  const int C = C0 + C1;
  if (((C0 | C1) & 3) == 0) {
    const GvDiv q = gv_div(C >> 2);
    const int q0 = C0 >> 2;
    for (int i = tid; i < nrow * q.d; i += GV_NT) {
      const int row = q.div(i), c4 = i - row * q.d;
      gv_cp16(dst + row * C + c4 * 4, c4 < q0 ? g0 + row * C0 + c4 * 4 : g1 + row * C1 + (c4 - q0) * 4);
    }
  } else {
    for (int i = tid; i < nrow * C; i += GV_NT) {
      const int row = i / C, c = i % C;
      dst[i] = c < C0 ? __ldg(g0 + row * C0 + c) : __ldg(g1 + row * C1 + (c - C0));
    }
  }
}

// Sum N per-lane partial vectors over the warp with a transposing butterfly (N/2 + N/4 + ... shuffles instead of 5 N):
// afterwards v[0] of lane l is the warp total of row gv_row<N>(l), replicated over the lanes that share that row.
template <int N>
__device__ __forceinline__ void gv_reduce(float* v, int lane, int off) {
  if constexpr (N > 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const float send = up ? v[i] : v[i + N / 2];
      const float keep = up ? v[i + N / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    gv_reduce<N / 2>(v, lane, off >> 1);
  } else {
    for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  }
}
template <int N>
__device__ __forceinline__ int gv_row(int lane) {
  int r = 0, off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1)
    if (lane & off) r += n / 2;
  return r;
}

// partial dot products of CW adjacent channels with RT output rows over K slice `ks` of `ns` (K = taps x C_in, flat; float4
// granules are dealt round-robin to the ns x 32 lanes working on the channels).  With CW = 2 every activation granule read
// from shared memory feeds two channels.  acc[cw * RT + r]; afterwards lane l holds in acc[0] the warp total of the
// flattened index gv_row<CW * RT>(l).
template <int RT, int CW, bool VEC>
__device__ __forceinline__ void gv_dot(const float* __restrict__ wcol, int wstride, const float* __restrict__ X, int Keff, int Cin,
                                       const int (*src)[GV_MAXL], int lane, int ks, int ns, float (&acc)[CW * RT]) {
#pragma unroll
  for (int r = 0; r < CW * RT; ++r) acc[r] = 0.f;
  const GvDiv dc = gv_div(Cin);
  if (VEC) {
#pragma unroll 2
    for (int k = (ks * 32 + lane) * 4; k < Keff; k += 128 * ns) {
      const int jj = dc.div(k), ci = k - jj * Cin;          // C_in % 4 == 0: a float4 never straddles two taps
      float4 w[CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) w[c] = *reinterpret_cast<const float4*>(wcol + c * wstride + k);
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(X + src[jj][r] + ci);
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[c * RT + r] += fmaf(w[c].x, x.x, w[c].y * x.y) + fmaf(w[c].z, x.z, w[c].w * x.w);
      }
    }
  } else {
    for (int k = ks * 32 + lane; k < Keff; k += 32 * ns) {
      const int jj = dc.div(k), ci = k - jj * Cin;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float x = X[src[jj][r] + ci];
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[c * RT + r] = fmaf(wcol[c * wstride + k], x, acc[c * RT + r]);
      }
    }
  }
  gv_reduce<CW * RT>(acc, lane, 16);
}

// x / n for a count that is (almost always) a power of two: the exact reciprocal built from the exponent instead of an IEEE division on the
// critical path between the dot product and the epilogue (same bits as the division)
__device__ __forceinline__ float gv_div_count(float x, int n) {
  return (n & (n - 1)) == 0 ? x * __uint_as_float((uint32_t)(127 - (31 - __clz(n))) << 23) : x / (float)n;
}

__host__ __device__ inline int gv_pad4(int n) { return (n + 7) & ~7; }   // floats; name kept: regions are padded to 32 bytes (see the carve-up)

// NC = channels per CTA (power of two <= 8), cls = CTAs per cluster (= GroupNorm group size / NC, or 1)
// CW = output channels per warp (2 when the CTA owns >= 2 channels: activation reads from shared memory are shared)
template <int RT, int CW>
__global__ void __launch_bounds__(GV_NT) conv_gemv_kernel(ConvArgs a, int NC, int cls, int trace_id) {
  extern __shared__ __align__(16) float dyn[];
  __shared__ GvStatic st;
  pdl_launch_dependents();       // let the following layers start streaming their weights right away
  GV_T(6); GV_T(0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (cls > 1) {
    if (tid == 0) gv_bar_init_expect(&st.bar, 8u * cls);   // this CTA will receive one (mean, M2) pair from every CTA of its cluster
    gv_cluster_arrive();         // "my mbarrier is initialised and I am running": peers wait for this before they send
  }
  const int col0 = blockIdx.x * NC, b = blockIdx.y;
  const int Cin = a.C0 + a.C1, RCin = a.RC0 + a.RC1;
  const int ntaps = a.jmax - a.jmin + 1;
  const int Keff = ntaps * Cin;
  const int L = a.Lout, Lin = a.Lin;
  const int nc = min(NC, a.Cout - col0);
  const int lnc = 31 - __clz(NC);                     // NC is a power of two
  const int lng = lnc - (CW == 2 ? 1 : 0);            // log2(channel groups of CW channels)
  const int ns = GV_NW >> lng;                        // K slices per channel group
  const bool vecW = (Cin & 3) == 0, vecR = (RCin & 3) == 0;
  // dynamic shared memory carve-up.  Every region is 32-BYTE aligned: a 16-byte cp.async whose shared-memory destination is 16 but
  // not 32 bytes aligned makes L2 return every 32-byte sector TWICE (measured: scripts/lts_bytes_bench.cu, profiles/r02_lts_bytes_bench.csv;
  // the dynamic block starts after the 3024-byte static block, which is what doubled lts__t_bytes of this path in round 1).
  // (pointer arithmetic on the __shared__ array, not a uintptr_t round trip: after an integer cast nvcc loses the address space and the dot-product
  //  loop reads its operands with generic LD.E.128 instead of LDS.128)
  float* Wsm = dyn + (((128u - ((unsigned)__cvta_generic_to_shared(dyn) & 127u)) & 127u) >> 2);   // [NC][Keff]
  float* RWsm = Wsm + gv_pad4(NC * Keff);             // [NC][RCin]
  float* X = RWsm + gv_pad4(NC * RCin);               // [Lin + 1][Cin], last row zero (padding taps read it)
  float* RX = X + gv_pad4((Lin + 1) * Cin);           // [L][RCin]

  // ---- everything that does not depend on the previous layer ----
  for (int w = 0; w < nc; ++w)
    gv_copy(Wsm + w * Keff, a.Wk + ((size_t)(col0 + w) * a.taps + a.jmin) * Cin, Keff, vecW, tid);
  if (a.resWk)
    for (int w = 0; w < nc; ++w) gv_copy(RWsm + w * RCin, a.resWk + (size_t)(col0 + w) * RCin, RCin, vecR, tid);
  if (tid < ntaps * GV_MAXL) {   // source-row table: which input row feeds (tap, output row)
    const int jj = tid / GV_MAXL, l = tid % GV_MAXL, j = a.jmin + jj;
    int pos;
    if (!a.transposed) pos = l * a.stride + j - a.pad;
    else { const int num = l + a.pad - j; pos = (num >= 0 && num % a.stride == 0) ? num / a.stride : -1; }
    st.src[jj][l] = (l < L && pos >= 0 && pos < Lin) ? pos * Cin : Lin * Cin;
  }
  if (tid < GV_MAXL) st.idsrc[0][tid] = (tid < L ? tid : 0) * RCin;
  for (int i = tid; i < Cin; i += GV_NT) X[Lin * Cin + i] = 0.f;
  // this thread's epilogue element (row er, channel ec) and its layer constants
  const int er = tid >> lnc, ew = tid & (NC - 1), ec = col0 + ew;
  const bool eown = tid < L * NC;                     // owns a (row, channel) slot of O
  const bool eact = eown && ew < nc;
  const bool gn = a.gn_gamma != nullptr;
  float e_gamma = 1.f, e_beta = 0.f, e_add = 0.f;
  if (eact && gn) { e_gamma = __ldg(a.gn_gamma + ec); e_beta = __ldg(a.gn_beta + ec); }
  const float e_bias = (eact && a.bias) ? __ldg(a.bias + ec) : 0.f;
  const float e_rb = (eact && a.resWk) ? __ldg(a.resB + ec) : 0.f;
  if (cls > 1) gv_cluster_wait();   // every peer is running and has initialised its mbarrier (off the dependency chain: before the wait)
  GV_T(1);
  pdl_wait();                    // inputs are produced by the preceding kernel
  GV_T(2);
  // ---- this sample's input activations ----
  gv_load_rows(X, a.x0 + (size_t)(a.x0_period > 0 ? b % a.x0_period : b) * Lin * a.C0, a.C0,
               a.C1 ? a.x1 + (size_t)b * Lin * a.C1 : nullptr, a.C1, Lin, tid);
  if (a.resWk)
    gv_load_rows(RX, a.rx0 + (size_t)(a.rx0_period > 0 ? b % a.rx0_period : b) * L * a.RC0, a.RC0,
                 a.RC1 ? a.rx1 + (size_t)b * L * a.RC1 : nullptr, a.RC1, L, tid);
  if (eact) {                    // additive epilogue terms: in flight while the dot products run
    if (a.temb) e_add += __ldg(a.temb + (size_t)b * a.temb_stride + ec);
    if (a.temb2) e_add += __ldg(a.temb2 + ec);
    if (a.res_id) e_add += __ldg(a.res_id + ((size_t)b * L + er) * a.Cout + ec);
  }
  gv_cp_wait_all();
  __syncthreads();
  GV_T(3);

  // ---- one warp per (K slice, group of CW output channels) ----
  {
    constexpr int NV = CW * RT;
    const int ch0 = (warp & ((NC >> (CW == 2 ? 1 : 0)) - 1)) * CW, ks = warp >> lng;
    const int f = gv_row<NV>(lane);                    // flattened (channel-in-group, row) this lane ends up holding
    const int ch = ch0 + f / RT, myrow = f % RT;
    const bool writer = (lane & (32 / NV - 1)) == 0 && myrow < L;
    float acc[NV];
    if (ch0 < nc) {                                     // nc == NC whenever CW == 2 (checked by the launcher)
      if (vecW) gv_dot<RT, CW, true>(Wsm + ch0 * Keff, Keff, X, Keff, Cin, st.src, lane, ks, ns, acc);
      else gv_dot<RT, CW, false>(Wsm + ch0 * Keff, Keff, X, Keff, Cin, st.src, lane, ks, ns, acc);
      if (writer) st.P[(ks * L + myrow) * NC + ch] = acc[0];
      if (a.resWk) {
        if (vecR) gv_dot<RT, CW, true>(RWsm + ch0 * RCin, RCin, RX, RCin, RCin, st.idsrc, lane, ks, ns, acc);
        else gv_dot<RT, CW, false>(RWsm + ch0 * RCin, RCin, RX, RCin, RCin, st.idsrc, lane, ks, ns, acc);
        if (writer) st.RP[(ks * L + myrow) * NC + ch] = acc[0];
      }
    } else if (writer) {
      st.P[(ks * L + myrow) * NC + ch] = 0.f;
      st.RP[(ks * L + myrow) * NC + ch] = 0.f;
    }
  }
  __syncthreads();
  GV_T(4);
  // combine the K slices in a fixed order (+ bias): done by the thread that owns (row, channel) in the epilogue
  float e_val = 0.f, e_res = 0.f;
  if (eown) {
    float v = st.P[tid];
    for (int k = 1; k < ns; ++k) v += st.P[k * L * NC + tid];
    e_val = ew < nc ? v + e_bias : 0.f;
    if (a.resWk) {
      float rv = st.RP[tid];
      for (int k = 1; k < ns; ++k) rv += st.RP[k * L * NC + tid];
      e_res = rv + e_rb;
    }
  }

  // ---- GroupNorm statistics: this CTA's NC channels x L rows are one part of the sample's group ----
  const int ne = NC * L;           // <= 128: the owners are the first ne threads
  float mu = 0.f, m2 = 0.f;
  if (gn) {
    if (eown) st.O[tid] = e_val;
    __syncthreads();
    if (warp == 0) {
      float v[4];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i] = lane + 32 * i < ne ? st.O[lane + 32 * i] : 0.f; s += v[i]; }
#pragma unroll
      for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
      const float mean = gv_div_count(s, ne);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) if (lane + 32 * i < ne) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
      for (int k = 16; k > 0; k >>= 1) q += __shfl_xor_sync(0xffffffffu, q, k);
      if (cls > 1) {
        if (lane < cls)                                     // publish this part's (mean, M2) to every CTA of the cluster
          gv_st_async(&st.cx[blockIdx.x % cls], &st.bar, (uint32_t)lane, mean, q);
      } else if (lane == 0) {
        st.mean = mean; st.m2 = q;
      }
    }
    if (cls > 1) {
      gv_bar_wait(&st.bar);                                 // all cls parts have landed in this CTA's cx (no CTA exits before that)
      float ms = 0.f;                                       // merge equal-sized parts in a fixed order (parallel-variance formula)
      for (int p = 0; p < cls; ++p) ms += st.cx[p].x;
      mu = gv_div_count(ms, cls);
      for (int p = 0; p < cls; ++p) { const float d = st.cx[p].x - mu; m2 += st.cx[p].y + (float)ne * d * d; }
    } else {
      __syncthreads();
      mu = st.mean; m2 = st.m2;
    }
  }
  GV_T(5);

  // ---- epilogue: one thread per (row, channel) ----
  if (eact) {
    float v = e_val;
    if (gn) {
      const float rstd = 1.0f / sqrtf(gv_div_count(m2, a.cg * L) + 1e-5f);
      v = mish_f((v - mu) * rstd * e_gamma + e_beta);
    }
    a.out[((size_t)b * L + er) * a.Cout + ec] = v + e_add + e_res;
  }
  GV_T(7);
}

// channels per CTA: aim at ~64 CTAs per sample, GroupNorm groups of at most 8 CTAs
static int gv_pick_nc(const ConvArgs& a) {
  int nc = 1;
  while (nc < GV_MAXNC && (a.Cout + nc - 1) / nc > 64) nc *= 2;
  if (a.gn_gamma) while (nc < GV_MAXNC && a.cg / nc > 8) nc *= 2;
  return nc;
}

static size_t gv_smem_floats(const ConvArgs& a, int nc) {
  const int Cin = a.C0 + a.C1, RCin = a.RC0 + a.RC1, Keff = (a.jmax - a.jmin + 1) * Cin;
  return (size_t)gv_pad4(nc * Keff) + gv_pad4(nc * RCin) + gv_pad4((a.Lin + 1) * Cin) + gv_pad4(a.Lout * RCin) + 32 /* 128-byte base alignment */;
}

template <int RT, int CW>
static int launch_rt(const ConvArgs& a, int nc, int cls, size_t smem, cudaStream_t s) {
  static size_t configured[64] = {};     // per device: function attributes belong to the device's context
  int dev = 0;
  B2P_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    B2P_CUDA_TRY(cudaFuncSetAttribute(conv_gemv_kernel<RT, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  prefer_max_smem_carveout((const void*)conv_gemv_kernel<RT, CW>);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((a.Cout + nc - 1) / nc, a.nrows / a.Lout); cfg.blockDim = dim3(GV_NT); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cls; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = cls > 1 ? 2 : 1;
  int trace_id = 0;
#ifdef B2P_GV_TRACE
  trace_id = gv_trace_launch++ & 8191;
#endif
  return (int)cudaLaunchKernelEx(&cfg, conv_gemv_kernel<RT, CW>, a, nc, cls, trace_id);
}

bool conv_gemv_applicable(const ConvArgs& a) {
  if (a.Lout > GV_MAXL || a.Lin > 2 * GV_MAXL || a.nrows % a.Lout != 0 || a.nrows / a.Lout > 65535) return false;
  if (!a.Wk || (a.resW && !a.resWk) || a.headW || a.out_hi || a.res_out || !a.out) return false;
  if (a.jmax - a.jmin + 1 > 5) return false;
  const int nc = gv_pick_nc(a);
  if (a.gn_gamma && (a.cg % nc != 0 || a.cg / nc > 8 || (a.Cout / nc) % (a.cg / nc) != 0)) return false;
  return gv_smem_floats(a, nc) * sizeof(float) <= 200 * 1024;
}

int launch_conv_gemv(const ConvArgs& a, cudaStream_t s) {
  if (!conv_gemv_applicable(a)) return B2P_ERR_INVALID_ARG;
  const int nc = gv_pick_nc(a);
  const size_t smem = gv_smem_floats(a, nc) * sizeof(float);
  const int cls = a.gn_gamma ? a.cg / nc : 1;               // CTAs sharing one GroupNorm group
  // two channels per warp when the CTA owns >= 2 complete channels (CW * RT accumulators must fit a warp's butterfly)
  // Two channels per warp halve the activation re-reads from shared memory but cost 64 instead of 40 registers per thread: a win
  // while the launch has at most one CTA per SM (1-2 samples: -4 % / -10 % per iteration), a loss once two CTAs share an SM and
  // the following layers' early-resident CTAs no longer fit the register file (4 samples: +7 %).
  static int cw_env = -1;
  if (cw_env < 0) { const char* e = getenv("B2P_GV_CW"); cw_env = e ? atoi(e) : 0; }
  const int ctas = ((a.Cout + nc - 1) / nc) * (a.nrows / a.Lout);
  const bool two = nc >= 2 && a.Cout % nc == 0 && a.Lout <= GV_MAXL / 2 &&   // (partials: 16 * CW * L floats must fit P)
                   (cw_env == 2 || (cw_env == 0 && ctas <= 148));
  if (a.Lout <= 2) return two ? launch_rt<2, 2>(a, nc, cls, smem, s) : launch_rt<2, 1>(a, nc, cls, smem, s);
  if (a.Lout <= 4) return two ? launch_rt<4, 2>(a, nc, cls, smem, s) : launch_rt<4, 1>(a, nc, cls, smem, s);
  if (a.Lout <= 8) return two ? launch_rt<8, 2>(a, nc, cls, smem, s) : launch_rt<8, 1>(a, nc, cls, smem, s);
  return launch_rt<16, 1>(a, nc, cls, smem, s);
}

}  // namespace b2p
