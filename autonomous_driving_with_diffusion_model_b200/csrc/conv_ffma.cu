// fp32 CUDA-core implicit-GEMM conv with the whole Conv1dBlock / residual-block epilogue fused
// (K3 in SURVEY.md Appendix C; the fp32 parity path).  Replaces, per launch, the reference's
//   Conv1d -> Rearrange -> GroupNorm(8) -> Rearrange -> Mish            (modeling/helpers.py:95-112)
//   + time_mlp broadcast add / residual_conv (1x1) add                   (modeling/temporal.py:53-55)
//   + torch.cat skip concat as a second K-slab                           (modeling/temporal.py:227)
//   + final 1x1 head and the [B,C,L] -> [B,L,C] rearrange                (modeling/temporal.py:233-245)
// Layout: channels-last rows (sample, position) x channels, so the trajectory tensor [B,H,D] is consumed and
// produced as is.  Tile: TM rows (whole samples) x 64 output channels (whole GroupNorm groups) per CTA.
#include <cuda_bf16.h>

#include "common.cuh"

namespace b2p {

// x / d, x % d for a runtime d that is a power of two for every layer of the model (shift / mask), any d otherwise
struct FastDiv {
  int d, sh;
  __device__ __forceinline__ int div(int x) const { return sh >= 0 ? x >> sh : x / d; }
  __device__ __forceinline__ int mod(int x) const { return sh >= 0 ? x & (d - 1) : x % d; }
};
__device__ __forceinline__ FastDiv fast_div(int d) {
  FastDiv r;
  r.d = d; r.sh = (d > 0 && (d & (d - 1)) == 0) ? 31 - __clz(d) : -1;
  return r;
}

constexpr int TN = 64;     // output channels per CTA
constexpr int KC = 16;     // K chunk
constexpr int NT = 256;    // threads
constexpr int APAD = 4;

template <int TM>
struct Smem {
  float A[2][TM][KC + APAD];
  float W[2][KC][TN];
  float C[TM][TN + 1];
  float mean[TM];   // one (sample, group) pair per entry, pairs <= TM*TN/(L*cg) <= TM (L*cg >= 64)
  float rstd[TM];
};

// position of tap j for output position l; returns -1 when the tap does not contribute
__device__ __forceinline__ int tap_pos(int l, int j, int stride, int pad, int transposed, int Lin) {
  int pos;
  if (!transposed) {
    pos = l * stride + j - pad;
  } else {
    int num = l + pad - j;
    if (num < 0 || (num % stride) != 0) return -1;
    pos = num / stride;
  }
  return (pos >= 0 && pos < Lin) ? pos : -1;
}

// One GEMM phase: acc[RM][4] += A(rows, K) * W(K, 64-col slice).  VEC: channel counts are multiples of 16.
template <int TM, bool VEC>
__device__ __forceinline__ void gemm_phase(Smem<TM>& sm, float (&acc)[TM / 16][4], const float* __restrict__ x0, int x0_period,
                                           const float* __restrict__ x1, int C0, int C1, int Lin, int Lout, int log2Lout,
                                           int nrows, int taps_lo, int taps_hi, int stride, int pad, int transposed,
                                           const float* __restrict__ W, int Cout, int row0, int col0) {
  constexpr int RM = TM / 16;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int Cin = C0 + C1;
  // loader roles
  const int arow = tid >> 2, akq = (tid & 3) * 4;      // A: row, 4 consecutive k
  const bool a_active = arow < TM;
  const int wk = tid >> 4, wc = (tid & 15) * 4;        // W: k row, 4 consecutive cols
  const int grow = row0 + arow;
  const int ab = grow >> log2Lout, al = grow & (Lout - 1);
  const int ab0 = x0_period > 0 ? ab % x0_period : ab;   // sample index into x0
  const bool row_ok = a_active && grow < nrows;

  int nchunks;
  if (VEC) nchunks = (taps_hi - taps_lo + 1) * (Cin / KC);
  else nchunks = ((taps_hi - taps_lo + 1) * Cin + KC - 1) / KC;

  const FastDiv dpt = fast_div(VEC ? Cin / KC : 1);   // chunks per tap
  float4 areg = make_float4(0.f, 0.f, 0.f, 0.f), wreg;
  auto load_chunk = [&](int ch) {
    if (VEC) {
      int j = taps_lo + dpt.div(ch);
      int c = dpt.mod(ch) * KC;
      areg = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) {
        int pos = tap_pos(al, j, stride, pad, transposed, Lin);
        if (pos >= 0) {
          int cc = c + akq;
          const float* src = (cc < C0) ? x0 + ((size_t)(ab0 * Lin + pos) * C0 + cc) : x1 + ((size_t)(ab * Lin + pos) * C1 + (cc - C0));
          areg = __ldg(reinterpret_cast<const float4*>(src));
        }
      }
      wreg = __ldg(reinterpret_cast<const float4*>(W + ((size_t)(j * Cin + c + wk) * Cout + col0 + wc)));
    } else {
      // flattened k = (j - taps_lo)*Cin + c  (first layer, Cin = transition_dim = 7)
      float v[4];
      int ktot = (taps_hi - taps_lo + 1) * Cin;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int kk = ch * KC + akq + i;
        v[i] = 0.f;
        if (row_ok && kk < ktot) {
          int j = taps_lo + kk / Cin, c = kk % Cin;
          int pos = tap_pos(al, j, stride, pad, transposed, Lin);
          if (pos >= 0) v[i] = (c < C0) ? __ldg(x0 + ((size_t)(ab0 * Lin + pos) * C0 + c)) : __ldg(x1 + ((size_t)(ab * Lin + pos) * C1 + (c - C0)));
        }
      }
      areg = make_float4(v[0], v[1], v[2], v[3]);
      int kk = ch * KC + wk;
      wreg = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < ktot) wreg = __ldg(reinterpret_cast<const float4*>(W + ((size_t)(taps_lo * Cin + kk) * Cout + col0 + wc)));
    }
  };
  auto store_chunk = [&](int buf) {
    if (a_active) *reinterpret_cast<float4*>(&sm.A[buf][arow][akq]) = areg;
    *reinterpret_cast<float4*>(&sm.W[buf][wk][wc]) = wreg;
  };

  if (nchunks <= 0) return;
  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    int buf = ch & 1;
    if (ch + 1 < nchunks) load_chunk(ch + 1);
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 a4[RM];
#pragma unroll
      for (int r = 0; r < RM; ++r) a4[r] = *reinterpret_cast<const float4*>(&sm.A[buf][ty * RM + r][k4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 w4 = *reinterpret_cast<const float4*>(&sm.W[buf][k4 + kk][tx * 4]);
#pragma unroll
        for (int r = 0; r < RM; ++r) {
          float av = kk == 0 ? a4[r].x : kk == 1 ? a4[r].y : kk == 2 ? a4[r].z : a4[r].w;
          acc[r][0] = fmaf(av, w4.x, acc[r][0]);
          acc[r][1] = fmaf(av, w4.y, acc[r][1]);
          acc[r][2] = fmaf(av, w4.z, acc[r][2]);
          acc[r][3] = fmaf(av, w4.w, acc[r][3]);
        }
      }
    }
    if (ch + 1 < nchunks) store_chunk(buf ^ 1);
    __syncthreads();
  }
}

template <int TM, bool VEC>
__global__ void __launch_bounds__(NT) conv_ffma_kernel(ConvArgs a) {
  constexpr int RM = TM / 16;
  __shared__ __align__(16) Smem<TM> sm;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * TM, col0 = blockIdx.y * TN;
  pdl_wait();                 // inputs come from the preceding kernel
  pdl_launch_dependents();

  float acc[RM][4];
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  gemm_phase<TM, VEC>(sm, acc, a.x0, a.x0_period, a.x1, a.C0, a.C1, a.Lin, a.Lout, a.log2Lout, a.nrows, a.jmin, a.jmax, a.stride, a.pad,
                      a.transposed, a.W, a.Cout, row0, col0);
  float racc[RM][4];
  if (a.resW) {  // residual 1x1 conv of the block input (modeling/temporal.py:40-44)
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) racc[r][c] = 0.f;
    if ((a.RC0 % KC) == 0 && (a.RC1 % KC) == 0)
      gemm_phase<TM, true>(sm, racc, a.rx0, a.rx0_period, a.rx1, a.RC0, a.RC1, a.Lout, a.Lout, a.log2Lout, a.nrows, 0, 0, 1, 0, 0, a.resW, a.Cout, row0, col0);
    else
      gemm_phase<TM, false>(sm, racc, a.rx0, a.rx0_period, a.rx1, a.RC0, a.RC1, a.Lout, a.Lout, a.log2Lout, a.nrows, 0, 0, 1, 0, 0, a.resW, a.Cout, row0, col0);
  }

  const int gc = col0 + tx * 4;  // global column of this thread's 4 outputs
  float4 bias4 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + gc)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < RM; ++r) { acc[r][0] += bias4.x; acc[r][1] += bias4.y; acc[r][2] += bias4.z; acc[r][3] += bias4.w; }

  if (a.gn_gamma) {
    // ---- GroupNorm(8) over (cg channels x Lout positions) per sample, biased variance, eps 1e-5 ----
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) sm.C[ty * RM + r][tx * 4 + c] = acc[r][c];
    __syncthreads();
    const int L = a.Lout, cg = a.cg;
    const int gpt = TN / cg;                 // groups per tile column block
    const int spt = TM / L;                  // samples per tile
    const int npairs = spt * gpt;
    const int ne = L * cg;
    const int warp = tid >> 5, lane = tid & 31;
    const FastDiv dcg = fast_div(cg), dgpt = fast_div(gpt), dL = fast_div(L);
    for (int p = warp; p < npairs; p += NT / 32) {
      int s = dgpt.div(p), g = dgpt.mod(p);
      float sum = 0.f;
      for (int e = lane; e < ne; e += 32) sum += sm.C[s * L + dcg.div(e)][g * cg + dcg.mod(e)];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      float mean = sum / (float)ne;
      float sq = 0.f;
      for (int e = lane; e < ne; e += 32) { float d = sm.C[s * L + dcg.div(e)][g * cg + dcg.mod(e)] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) { sm.mean[p] = mean; sm.rstd[p] = 1.0f / sqrtf(sq / (float)ne + 1e-5f); }
    }
    __syncthreads();
    float4 gm = __ldg(reinterpret_cast<const float4*>(a.gn_gamma + gc));
    float4 bt = __ldg(reinterpret_cast<const float4*>(a.gn_beta + gc));
    const int g = dcg.div(tx * 4);
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      int lr = ty * RM + r;
      int p = dL.div(lr) * gpt + g;
      float mean = sm.mean[p], rstd = sm.rstd[p];
      acc[r][0] = mish_f((acc[r][0] - mean) * rstd * gm.x + bt.x);
      acc[r][1] = mish_f((acc[r][1] - mean) * rstd * gm.y + bt.y);
      acc[r][2] = mish_f((acc[r][2] - mean) * rstd * gm.z + bt.z);
      acc[r][3] = mish_f((acc[r][3] - mean) * rstd * gm.w + bt.w);
    }
  }

  float4 rb4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.resW) rb4 = __ldg(reinterpret_cast<const float4*>(a.resB + gc));
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    int lr = ty * RM + r;
    int grow = row0 + lr;
    bool ok = grow < a.nrows;
    if (a.temb && ok) {
      float4 t4 = __ldg(reinterpret_cast<const float4*>(a.temb + (size_t)(grow >> a.log2Lout) * a.temb_stride + gc));
      acc[r][0] += t4.x; acc[r][1] += t4.y; acc[r][2] += t4.z; acc[r][3] += t4.w;
    }
    if (a.temb2) {
      float4 t4 = __ldg(reinterpret_cast<const float4*>(a.temb2 + gc));
      acc[r][0] += t4.x; acc[r][1] += t4.y; acc[r][2] += t4.z; acc[r][3] += t4.w;
    }
    if (a.res_id && ok) {
      float4 q = __ldg(reinterpret_cast<const float4*>(a.res_id + (size_t)grow * a.Cout + gc));
      acc[r][0] += q.x; acc[r][1] += q.y; acc[r][2] += q.z; acc[r][3] += q.w;
    }
    if (a.resW) {
      if (a.res_out) {   // emit the 1x1 projection on its own (consumed as a residual by a later tensor-core launch)
        if (ok) *reinterpret_cast<float4*>(a.res_out + (size_t)grow * a.Cout + gc) =
            make_float4(racc[r][0] + rb4.x, racc[r][1] + rb4.y, racc[r][2] + rb4.z, racc[r][3] + rb4.w);
      } else {
        acc[r][0] += racc[r][0] + rb4.x; acc[r][1] += racc[r][1] + rb4.y;
        acc[r][2] += racc[r][2] + rb4.z; acc[r][3] += racc[r][3] + rb4.w;
      }
    }
    if (a.out && ok)
      *reinterpret_cast<float4*>(a.out + (size_t)grow * a.Cout + gc) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    if (a.out_hi && ok) {   // bf16 hi (+lo) copy for the tensor-core layers that follow
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { h[c] = __float2bfloat16_rn(acc[r][c]); l[c] = __float2bfloat16_rn(acc[r][c] - __bfloat162float(h[c])); }
      *reinterpret_cast<uint2*>(a.out_hi + (size_t)grow * a.Cout + gc) = *reinterpret_cast<uint2*>(h);
      if (a.out_lo) *reinterpret_cast<uint2*>(a.out_lo + (size_t)grow * a.Cout + gc) = *reinterpret_cast<uint2*>(l);
    }
  }

  if (a.headW) {  // fused 1x1 head over the 64 channels of each row (Cout == 64, gridDim.y == 1)
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) sm.C[ty * RM + r][tx * 4 + c] = acc[r][c];
    __syncthreads();
    const int hd = a.head_dim;
    for (int i = tid; i < TM * hd; i += NT) {
      int lr = i / hd, d = i % hd;
      int grow = row0 + lr;
      if (grow >= a.nrows) continue;
      float s = __ldg(a.headB + d);
#pragma unroll 8
      for (int c = 0; c < TN; ++c) s = fmaf(sm.C[lr][c], __ldg(a.headW + c * hd + d), s);
      a.head_out[(size_t)grow * hd + d] = s;
    }
  }
}

template <int TM>
static int launch_tm(const ConvArgs& a, bool vec, cudaStream_t s) {
  dim3 grid((a.nrows + TM - 1) / TM, a.Cout / TN);
  // plain (fully serialised) launches: these grids fill the machine, so early-resident dependents only take SM slots away
  prefer_max_smem_carveout(vec ? (const void*)conv_ffma_kernel<TM, true> : (const void*)conv_ffma_kernel<TM, false>);
  if (vec) conv_ffma_kernel<TM, true><<<grid, NT, 0, s>>>(a);
  else conv_ffma_kernel<TM, false><<<grid, NT, 0, s>>>(a);
  return (int)cudaGetLastError();
}

int launch_conv_ffma(const ConvArgs& a, cudaStream_t s) {
  if (a.Cout % TN != 0 || a.nrows <= 0) return B2P_ERR_INVALID_ARG;
  if (a.headW && a.Cout != TN) return B2P_ERR_INVALID_ARG;
  bool vec = (a.C0 % KC == 0) && (a.C1 % KC == 0);
  // tile rows: whole samples, and with GroupNorm at most TM (sample, group) pairs per tile
  int min_tm = a.Lout < 16 ? 16 : a.Lout;
  if (min_tm > 64) return B2P_ERR_INVALID_ARG;
  if (a.gn_gamma && (a.Lout * a.cg < TN || TN % a.cg != 0)) return B2P_ERR_INVALID_ARG;
  // pick the largest tile that still gives >= 148 CTAs, else the smallest legal one
  int ntile_n = a.Cout / TN;
  int tm = 64;
  while (tm > min_tm && ((a.nrows + tm - 1) / tm) * ntile_n < 148) tm >>= 1;
  if (tm == 64) return launch_tm<64>(a, vec, s);
  if (tm == 32) return launch_tm<32>(a, vec, s);
  return launch_tm<16>(a, vec, s);
}

}  // namespace b2p
