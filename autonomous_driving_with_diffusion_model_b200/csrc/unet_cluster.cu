// EXPERIMENTAL, OPT-IN (B2P_CLUSTER_EVAL=1), NOT ON ANY DEFAULT PATH.
// Status (round 1, profiles/r01_cluster_eval_experiment.txt): CORRECT on B200 (forward within 7.6e-7 of the GEMV path, DDIM-10
// plan within 8.4e-7; program builder + algorithm also checked on the CPU by tests/test_cluster_program.py) but NOT YET FASTER:
// 224 us per denoising iteration against 172 us.  Timing bisect per evaluation: dot products 43 us, DSMEM exchange + cluster
// barrier 42 us, waiting for weights 13 us, and a 127 us skeleton (tables, ~8 block barriers and one exposed L2 round trip for
// the per-channel constants per layer, instruction fetch of four inlined variants) that the estimate below did not foresee.
//
// Whole-denoiser evaluation (TemporalMapUnet.forward, modeling/temporal.py:197-245, NO_GUIDANCE) of ONE trajectory by ONE
// thread-block cluster of 16 CTAs in ONE launch, for closed-loop planning (one trajectory per tick,
// e2e_driving/diffusion_agent.py:179).  Today's small-batch path (conv_gemv.cu) is a chain of 42 dependent launches per
// evaluation at ~3.9 us each (164 us per denoising iteration, profiles/r01_gemv_stage_trace_b1.txt): activation fetch
// from L2 0.75 + dot 0.7-1.1 + GroupNorm exchange 1.3 + dependency release 1.1.  Here:
//   * every CTA keeps a full copy of every live activation tensor of the trajectory in shared memory (<= 1024 floats each,
//     8 slots), so no activation ever goes through L2 and no launch boundary exists between layers;
//   * CTA r computes output channels [r*nc, (r+1)*nc) of each layer (nc = C_out/16) from its own slice of the layer's
//     K-major fp32 weights, which arrive through a 4 x 32 KB ring of 1-D bulk copies (cp.async.bulk + mbarrier) from a
//     per-CTA contiguous stream: the stream runs ahead across layer boundaries, only bounded by the ring;
//   * raw conv outputs (+bias) are pushed to all 16 CTAs with DSMEM stores, one cluster barrier per layer, and EVERY CTA
//     normalises the whole tensor itself (GroupNorm statistics + Mish + time term + residual on <= 1024 elements: two per
//     thread), so there is no statistics exchange at all ("normalisation at the consumer").
// Design estimate (not met, see status): ~1.2 us per layer (cluster barrier 0.25 + block barriers 0.4 + dot 0.3 + epilogue 0.2) => ~50 us per
// evaluation, with the 2.8 MB-per-CTA weight stream (at ~80 B/clk per SM: ~18 us) hidden underneath.
//
// Numerics: exact fp32, same formulas as conv_gemv.cu (two-pass GroupNorm statistics, mish_f, bias before statistics);
// summation order differs (K slices), so results agree with the other kernels to fp32 rounding, not bit for bit.
#include "common.cuh"
#include "unet_cluster.cuh"

namespace b2p {

namespace {

// ---- shared-memory carve-up (floats) ----
constexpr int UC_O_RING = 0;
constexpr int UC_O_SLOTS = UC_O_RING + UC_NSTAGE * UC_STAGE_FLOATS;
constexpr int UC_O_ZERO = UC_O_SLOTS + UC_NSLOT * UC_SLOT_FLOATS;    // 1024 zeros: rows that fall into the conv padding
constexpr int UC_O_RAW = UC_O_ZERO + UC_SLOT_FLOATS;                 // [parity 2][conv, residual][1024]
constexpr int UC_O_P = UC_O_RAW + 4 * UC_SLOT_FLOATS;                // K-slice partials of the conv      [slice][row][channel] <= 512
constexpr int UC_O_RP = UC_O_P + 512;                                // same for the residual 1x1 conv
constexpr int UC_O_O = UC_O_RP + 512;                                // this CTA's outputs: [64] conv + [64] residual
constexpr int UC_O_STAT = UC_O_O + 128;                              // mean[8], rstd[8]
constexpr int UC_O_HEAD = UC_O_STAT + 16;                            // head weights [head_dim][64] (<= 448) + bias at +448
constexpr int UC_O_SRC = UC_O_HEAD + 512;                            // int tables: src[2][5][16], idsrc[2][16]
constexpr int UC_O_PROG = UC_O_SRC + 192;                            // UcProgram copy (16-byte aligned: all offsets are multiples of 4)
static_assert(sizeof(UcProgram) % 16 == 0, "UcProgram is copied in 16-byte units");
constexpr size_t UC_SMEM_BYTES = (size_t)UC_O_PROG * 4 + ((sizeof(UcProgram) + 15) & ~(size_t)15) + 64;   // + mbarriers

#ifdef B2P_TC_TRACE   // developer tracing (trace build, scripts/cluster_trace.py): stage clocks of CTA 0 of trajectory 0, per layer
__device__ unsigned long long uc_trace[(UC_MAXOPS + 1) * 8];
#define UC_T(op, k)                                                                  \
  do {                                                                               \
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) uc_trace[(op) * 8 + (k)] = clock64(); \
  } while (0)
#else
#define UC_T(op, k)
#endif

__device__ __forceinline__ uint32_t uc_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void uc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void uc_st_remote(const float* local_dst, uint32_t peer, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(uc_saddr(local_dst)), "r"(peer));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
// 8-byte asynchronous store into CTA `peer` that also counts its bytes on THAT CTA's mbarrier (conv_gemv.cu's exchange): the
// receiver waits for its byte count, no cluster-wide barrier or fence is involved
__device__ __forceinline__ void uc_st_async(const float* local_dst, const unsigned long long* local_bar, uint32_t peer, float a, float b) {
  uint32_t rd, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rd) : "r"(uc_saddr(local_dst)), "r"(peer));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(uc_saddr(local_bar)), "r"(peer));
  const unsigned long long v = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(rd), "l"(v), "r"(rb) : "memory");
}
__device__ __forceinline__ void uc_bar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_saddr(bar)) : "memory");
}
__device__ __forceinline__ void uc_bar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "UCW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra UCD_%=;\n\t"
      "bra UCW_%=;\n\t"
      "UCD_%=:\n\t}" ::"r"(uc_saddr(bar)), "r"(parity) : "memory");
}
// one chunk of this CTA's weight stream -> ring stage, completion counted in bytes on the stage's mbarrier
__device__ __forceinline__ void uc_issue(const UcProgram* pg, int q, const float* stream, float* ring, unsigned long long* full) {
  if (q >= pg->n_chunks) return;
  const UcChunk& c = pg->chunks[q];
  const int stage = q % UC_NSTAGE;
  const uint32_t bar = uc_saddr(full + stage), dst = uc_saddr(ring + stage * UC_STAGE_FLOATS);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)c.bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(stream + c.off),
               "r"((uint32_t)c.bytes), "r"(bar)
               : "memory");
}

// warp sum of N per-lane partials with a transposing butterfly: afterwards v[0] of lane l is the warp total of flattened
// index uc_row<N>(l), replicated over the lanes that share it (same scheme as conv_gemv.cu)
template <int N>
__device__ __forceinline__ void uc_reduce(float* v, int lane, int off) {
  if constexpr (N > 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const float send = up ? v[i] : v[i + N / 2];
      const float keep = up ? v[i + N / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    uc_reduce<N / 2>(v, lane, off >> 1);
  } else {
    for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  }
}
template <int N>
__device__ __forceinline__ int uc_row(int lane) {
  int r = 0, off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1)
    if (lane & off) r += n / 2;
  return r;
}

// acc[c * RT + r] += sum over this warp's share of K range [k0, k0 + klen) of W[ch0 + c][k] * X[row(tap(k), r)][channel(k)]
// wst = stage + ch0 * kstride; tables t0 / t1: float offset (from sm) of the input row feeding (tap, output row) in source 0 / 1
template <int RT, bool VEC>
__device__ __forceinline__ void uc_dot(float (&acc)[2 * RT], const float* __restrict__ wst, int kstride, int k0, int klen, const float* sm,
                                       const int* __restrict__ t0, const int* __restrict__ t1, int C0, int Cin, int cin_shift, int lane, int ks,
                                       int ns) {
  if (VEC) {
    for (int kk = (ks * 32 + lane) * 4; kk < klen; kk += 128 * ns) {
      const int k = k0 + kk;
      const int jj = cin_shift >= 0 ? k >> cin_shift : k / Cin, ci = k - jj * Cin;
      const float4 w0 = *reinterpret_cast<const float4*>(wst + kk);
      const float4 w1 = *reinterpret_cast<const float4*>(wst + kstride + kk);
      const bool first = ci < C0;
      const int* tb = (first ? t0 : t1) + jj * UC_MAXL;
      const int cc = first ? ci : ci - C0;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(sm + tb[r] + cc);
        acc[r] += fmaf(w0.x, x.x, w0.y * x.y) + fmaf(w0.z, x.z, w0.w * x.w);
        acc[RT + r] += fmaf(w1.x, x.x, w1.y * x.y) + fmaf(w1.z, x.z, w1.w * x.w);
      }
    }
  } else {
    for (int kk = ks * 32 + lane; kk < klen; kk += 32 * ns) {
      const int k = k0 + kk;
      const int jj = cin_shift >= 0 ? k >> cin_shift : k / Cin, ci = k - jj * Cin;
      const float w0 = wst[kk], w1 = wst[kstride + kk];
      const bool first = ci < C0;
      const int* tb = (first ? t0 : t1) + jj * UC_MAXL;
      const int cc = first ? ci : ci - C0;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float x = sm[tb[r] + cc];
        acc[r] = fmaf(w0, x, acc[r]);
        acc[RT + r] = fmaf(w1, x, acc[RT + r]);
      }
    }
  }
}

// consume `nchunks` ring stages (chunks q, q+1, ...) of one weight matrix: acc += W_slice . X over each chunk's K range
template <int RT>
__device__ __forceinline__ void uc_consume(float (&acc)[2 * RT], const UcProgram* pg, int nchunks, int& q, float* sm, const float* stream,
                                           unsigned long long* full, const int* t0, const int* t1, int C0, int C1, int ch0, int lane, int ks,
                                           int ns, int tid, int dbg) {
  float* ring = sm + UC_O_RING;
  const int Cin = C0 + C1;
  const int sh = (Cin & (Cin - 1)) == 0 ? 31 - __clz(Cin) : -1;
  const bool vec = ((C0 | C1) & 3) == 0;
  for (int c = 0; c < nchunks; ++c, ++q) {
    const UcChunk& ck = pg->chunks[q];
    const int stage = q % UC_NSTAGE;
    if (!(dbg & 8)) uc_bar_wait(full + stage, (uint32_t)((q / UC_NSTAGE) & 1));
    const float* wst = ring + stage * UC_STAGE_FLOATS + ch0 * ck.kstride;
    if (!(dbg & 2)) {
      if (vec) uc_dot<RT, true>(acc, wst, ck.kstride, ck.k0, ck.klen, sm, t0, t1, C0, Cin, sh, lane, ks, ns);
      else uc_dot<RT, false>(acc, wst, ck.kstride, ck.k0, ck.klen, sm, t0, t1, C0, Cin, sh, lane, ks, ns);
    }
    __syncthreads();                        // every warp is done with the stage
    if (tid == 0 && !(dbg & 8)) uc_issue(pg, q + UC_NSTAGE, stream, ring, full);
  }
}

// dot-product phase of one op: conv chunks, then the residual 1x1 chunks, from the ring; K-slice partials to P / RP
template <int RT>
__device__ __forceinline__ void uc_layer_dot(const UcProgram* pg, const UcOp& o, float* sm, const float* stream, unsigned long long* full, int& q,
                                             int tid, int dbg) {
  constexpr int NV = 2 * RT;
  const int warp = tid >> 5, lane = tid & 31;
  const int ng = o.nc >> 1;                 // channel pairs owned by this CTA (power of two, <= 16)
  const int ns = (UC_NT / 32) / ng;         // K slices
  const int grp = warp & (ng - 1), ks = warp / ng;
  const int ch0 = grp * 2;
  const int* t0 = reinterpret_cast<const int*>(sm + UC_O_SRC);
  const int* t1 = t0 + 5 * UC_MAXL;
  const int* i0 = t1 + 5 * UC_MAXL;         // a 1x1 conv is the one-tap case: its "tap" tables are the identity-row tables
  const int* i1 = i0 + UC_MAXL;
  const int f = uc_row<NV>(lane);
  const int ch = ch0 + f / RT, myrow = f % RT;
  const bool writer = (lane & (32 / NV - 1)) == 0 && myrow < o.Lout;

  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  uc_consume<RT>(acc, pg, o.nchunks, q, sm, stream, full, t0, t1, o.C0, o.C1, ch0, lane, ks, ns, tid, dbg);
  uc_reduce<NV>(acc, lane, 16);
  if (writer) sm[UC_O_P + (ks * o.Lout + myrow) * o.nc + ch] = acc[0];
  if (o.rnchunks > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    uc_consume<RT>(acc, pg, o.rnchunks, q, sm, stream, full, i0, i1, o.RC0, o.RC1, ch0, lane, ks, ns, tid, dbg);
    uc_reduce<NV>(acc, lane, 16);
    if (writer) sm[UC_O_RP + (ks * o.Lout + myrow) * o.nc + ch] = acc[0];
  }
}

__global__ void __launch_bounds__(UC_NT, 1) unet_cluster_kernel(UcLaunch a) {
  extern __shared__ __align__(128) float sm[];
  pdl_launch_dependents();
  UC_T(UC_MAXOPS, 0);
  const int tid = threadIdx.x;
  const int rank = blockIdx.x;              // grid = (UC_CL, B), cluster = (UC_CL, 1, 1): rank in the cluster == blockIdx.x
  const int b = blockIdx.y;
  const int dbg = a.dbg;                    // developer timing bisect (results are wrong when any of bits 2/4/8 is set)
  UcProgram* pg = reinterpret_cast<UcProgram*>(sm + UC_O_PROG);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(pg) + ((sizeof(UcProgram) + 15) & ~(size_t)15));

  // ---- prologue: nothing here depends on the preceding kernel ----
  {
    const int4* src = reinterpret_cast<const int4*>(a.prog);
    int4* dst = reinterpret_cast<int4*>(pg);
    for (int i = tid; i < (int)(sizeof(UcProgram) / 16); i += UC_NT) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < UC_SLOT_FLOATS; i += UC_NT) sm[UC_O_ZERO + i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < UC_NSTAGE + 2; ++s) uc_bar_init(full + s);   // ring stages + the two receive barriers (raw buffer parity)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");   // the bulk copies (async proxy) signal these barriers
  }
  __syncthreads();
  const float* stream = a.stream + (size_t)rank * pg->stream_floats_per_cta;
  if (tid == 0 && !(dbg & 8))
    for (int s = 0; s < UC_NSTAGE; ++s) uc_issue(pg, s, stream, sm + UC_O_RING, full);
  for (int i = tid; i < pg->head_dim * 64; i += UC_NT) sm[UC_O_HEAD + i] = __ldg(a.pack + pg->headWk + i);
  if (tid < pg->head_dim) sm[UC_O_HEAD + 448 + tid] = __ldg(a.pack + pg->headB + tid);
  uc_cluster_sync();                        // every CTA of the cluster is running: its shared memory may be written remotely
  pdl_wait();                               // the trajectory and the time terms come from preceding kernels
  for (int i = tid; i < a.H * a.D; i += UC_NT) sm[UC_O_SLOTS + pg->x_slot * UC_SLOT_FLOATS + i] = __ldg(a.x + (size_t)b * a.H * a.D + i);
  __syncthreads();

  UC_T(UC_MAXOPS, 1);                       // prologue done (slot 0 of the last row = kernel start, below)
  int q = 0;                                // next chunk of the stream to consume
  for (int oi = 0; oi < pg->n_ops; ++oi) {
    UC_T(oi, 0);
    const UcOp& o = pg->ops[oi];
    const int par = oi & 1;
    float* raw = sm + UC_O_RAW + par * 2 * UC_SLOT_FLOATS;
    float* rraw = raw + UC_SLOT_FLOATS;
    const int n_items = o.Lout * o.nc;      // outputs this CTA produces (<= 64)
    const int n_el = o.Lout * o.Cout;       // elements of the layer output (<= 1024)
    const bool has_res = o.rnchunks > 0;
    // UNTESTED VARIANT (dbg bit 16, i.e. B2P_CLUSTER_EVAL=17): exchange by st.async + byte counting on the receiver's mbarrier
    // instead of remote stores + a cluster barrier.  The phase of rx[par] used two layers ago has completed (every thread of
    // this CTA waited on it), and no peer can send this layer's data before it has received this CTA's previous layer.
    const bool xasync = (dbg & 16) && !(dbg & 4);
    unsigned long long* rx = full + UC_NSTAGE;
    if (xasync && tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(uc_saddr(rx + par)), "r"((uint32_t)(n_el * 4 * (has_res ? 2 : 1)))
                   : "memory");

    // (a) input-row tables
    if (tid < 2 * 5 * UC_MAXL) {
      const int s = tid / (5 * UC_MAXL), jj = (tid / UC_MAXL) % 5, l = tid % UC_MAXL;
      const int slot = s == 0 ? o.in0 : o.in1, C = s == 0 ? o.C0 : o.C1;
      int off = UC_O_ZERO;
      if (slot >= 0 && jj < o.ntaps && l < o.Lout) {
        const int j = o.jmin + jj;
        int pos;
        if (!o.transposed) pos = l * o.stride + j - o.pad;
        else { const int num = l + o.pad - j; pos = (num >= 0 && num % o.stride == 0) ? num / o.stride : -1; }
        if (pos >= 0 && pos < o.Lin) off = UC_O_SLOTS + slot * UC_SLOT_FLOATS + pos * C;
      }
      reinterpret_cast<int*>(sm + UC_O_SRC)[tid] = off;
    } else if (tid < 2 * 5 * UC_MAXL + 2 * UC_MAXL) {
      const int t = tid - 2 * 5 * UC_MAXL, s = t / UC_MAXL, l = t % UC_MAXL;
      const int slot = s == 0 ? o.rin0 : o.rin1, C = s == 0 ? o.RC0 : o.RC1;
      reinterpret_cast<int*>(sm + UC_O_SRC)[tid] = (slot >= 0 && l < o.Lout) ? UC_O_SLOTS + slot * UC_SLOT_FLOATS + l * C : UC_O_ZERO;
    }
    // (b) constants of this thread's epilogue elements and of the output it produces: in flight during the dot products
    float e_gamma[2] = {1.f, 1.f}, e_beta[2] = {0.f, 0.f}, e_add[2] = {0.f, 0.f};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int e = tid + s * UC_NT;
      if (e < n_el) {
        const int ch = e & (o.Cout - 1);
        if (o.gn) { e_gamma[s] = __ldg(a.pack + o.gamma + ch); e_beta[s] = __ldg(a.pack + o.beta + ch); }
        if (o.temb_off >= 0) {
          e_add[s] = __ldg(a.temb + (size_t)b * a.temb_stride + o.temb_off + ch);
          if (a.temb2) e_add[s] += __ldg(a.temb2 + o.temb_off + ch);
        }
      }
    }
    float p_bias = 0.f, p_rb = 0.f;
    if (tid < n_items) {
      const int ch = rank * o.nc + (tid & (o.nc - 1));
      if (o.bias >= 0) p_bias = __ldg(a.pack + o.bias + ch);
      if (has_res && o.resB >= 0) p_rb = __ldg(a.pack + o.resB + ch);
    }
    __syncthreads();
    UC_T(oi, 1);

    // (c) dot products over this CTA's weight slice
    switch (o.Lout) {
      case 2: uc_layer_dot<2>(pg, o, sm, stream, full, q, tid, dbg); break;
      case 4: uc_layer_dot<4>(pg, o, sm, stream, full, q, tid, dbg); break;
      case 8: uc_layer_dot<8>(pg, o, sm, stream, full, q, tid, dbg); break;
      default: uc_layer_dot<16>(pg, o, sm, stream, full, q, tid, dbg); break;
    }
    __syncthreads();
    UC_T(oi, 2);

    // (d) combine the K slices in a fixed order (+ bias)
    if (tid < n_items) {
      const int ns = (UC_NT / 32) / (o.nc >> 1);
      float v = sm[UC_O_P + tid];
      for (int k = 1; k < ns; ++k) v += sm[UC_O_P + k * n_items + tid];
      sm[UC_O_O + tid] = v + p_bias;
      if (has_res) {
        float rv = sm[UC_O_RP + tid];
        for (int k = 1; k < ns; ++k) rv += sm[UC_O_RP + k * n_items + tid];
        sm[UC_O_O + 64 + tid] = rv + p_rb;
      }
    }
    __syncthreads();
    UC_T(oi, 3);

    // (e) push this CTA's outputs into the raw buffer of every CTA of the cluster (itself included)
    if (xasync) {
      const int n_pairs = n_items >> 1;       // nc is even: items (it, it + 1) are adjacent channels of one row
      for (int idx = tid; idx < n_pairs * UC_CL; idx += UC_NT) {
        const int peer = idx / n_pairs, it = (idx - peer * n_pairs) * 2;
        const int r = it / o.nc, cl = it - r * o.nc;
        const int e = r * o.Cout + rank * o.nc + cl;
        uc_st_async(raw + e, rx + par, (uint32_t)peer, sm[UC_O_O + it], sm[UC_O_O + it + 1]);
        if (has_res) uc_st_async(rraw + e, rx + par, (uint32_t)peer, sm[UC_O_O + 64 + it], sm[UC_O_O + 64 + it + 1]);
      }
      uc_bar_wait(rx + par, (uint32_t)((oi >> 1) & 1));   // all 16 parts of this layer have landed here
    } else if (!(dbg & 4)) {
      for (int idx = tid; idx < n_items * UC_CL; idx += UC_NT) {
        const int peer = idx / n_items, it = idx - peer * n_items;
        const int r = it / o.nc, cl = it - r * o.nc;
        const int e = r * o.Cout + rank * o.nc + cl;
        uc_st_remote(raw + e, (uint32_t)peer, sm[UC_O_O + it]);
        if (has_res) uc_st_remote(rraw + e, (uint32_t)peer, sm[UC_O_O + 64 + it]);
      }
      // (f) one cluster barrier per layer: all parts of the layer output have landed everywhere
      uc_cluster_sync();
    } else {
      __syncthreads();
    }
    UC_T(oi, 4);

    // (g) GroupNorm statistics of the whole tensor, redundantly in every CTA: warp g owns group g
    if (o.gn) {
      const int warp = tid >> 5, lane = tid & 31;
      if (warp < 8) {
        const int cg = o.Cout >> 3, n = o.Lout * cg;   // 64 or 128 elements
        float v[4];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = lane + 32 * i;
          const int r = j / cg, c = j - r * cg;
          v[i] = j < n ? raw[r * o.Cout + warp * cg + c] : 0.f;
          s += v[i];
        }
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
        const float mean = s / (float)n;
        float qq = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (lane + 32 * i < n) { const float d = v[i] - mean; qq = fmaf(d, d, qq); }
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, k);
        if (lane == 0) { sm[UC_O_STAT + warp] = mean; sm[UC_O_STAT + 8 + warp] = 1.0f / sqrtf(qq / (float)n + 1e-5f); }
      }
      __syncthreads();
    }
    UC_T(oi, 5);

    // (h) epilogue on the whole tensor: normalise, Mish, time term, residual -> this layer's output slot
    float* out = sm + UC_O_SLOTS + o.out * UC_SLOT_FLOATS;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int e = tid + s * UC_NT;
      if (e < n_el) {
        float v = raw[e];
        if (o.gn) {
          const int g = (e & (o.Cout - 1)) / (o.Cout >> 3);
          v = mish_f((v - sm[UC_O_STAT + g]) * sm[UC_O_STAT + 8 + g] * e_gamma[s] + e_beta[s]);
        }
        float res = 0.f;
        if (has_res) res = rraw[e];
        else if (o.res_id >= 0) res = sm[UC_O_SLOTS + o.res_id * UC_SLOT_FLOATS + e];
        out[e] = v + e_add[s] + res;
      }
    }
    __syncthreads();
    UC_T(oi, 6);

    // (i) 1x1 head on the finished [L][64] tensor (final_conv.1): written once, by CTA 0
    if (o.head && rank == 0 && tid < o.Lout * pg->head_dim) {
      const int r = tid / pg->head_dim, j = tid - r * pg->head_dim;
      const float* xr = out + r * o.Cout;
      const float* w = sm + UC_O_HEAD + j * 64;
      float v = 0.f;
      for (int c = 0; c < 64; ++c) v = fmaf(xr[c], w[c], v);
      a.head_out[((size_t)b * o.Lout + r) * pg->head_dim + j] = v + sm[UC_O_HEAD + 448 + j];
    }
  }
}

}  // namespace

size_t uc_smem_bytes() { return UC_SMEM_BYTES; }

int launch_unet_cluster(const UcLaunch& a, cudaStream_t s) {
  static bool configured[64] = {};
  int dev = 0;
  B2P_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return B2P_ERR_INVALID_ARG;
  if (!configured[dev]) {
    B2P_CUDA_TRY(cudaFuncSetAttribute(unet_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UC_SMEM_BYTES));
    B2P_CUDA_TRY(cudaFuncSetAttribute(unet_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured[dev] = true;
  }
  prefer_max_smem_carveout((const void*)unet_cluster_kernel);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(UC_CL, a.B); cfg.blockDim = dim3(UC_NT); cfg.dynamicSmemBytes = UC_SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = UC_CL; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return (int)cudaLaunchKernelEx(&cfg, unet_cluster_kernel, a);
}

}  // namespace b2p

#ifdef B2P_TC_TRACE
// rows 0..n_ops-1: clocks at {layer start, tables built, dot products done, K slices combined, exchange + cluster barrier done,
// statistics done, epilogue done}; row UC_MAXOPS: {kernel start, prologue done}
extern "C" __attribute__((visibility("default"))) int b2p_debug_uc_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, b2p::uc_trace, sizeof(unsigned long long) * (b2p::UC_MAXOPS + 1) * 8);
}
#endif
