// TrajPredict (classifier-guidance state predictor) forward, and the whole classifier-guidance update with an
// ANALYTIC backward pass (K7 in SURVEY.md Appendix C).  One CTA per trajectory; every linear layer (forward and backward) runs on the
// tensor cores (mma.sync 3xTF32, see linear()); all activations of the 2-layer
// post-LN transformer (seq 15, d 64, 4 heads, SiLU FFN 256) stay in shared memory between forward and backward.
// Replaces: modeling/helpers.py:22-59 (TrajPredict.forward), interact.py:154-160 (state/model_output assembly),
//           control/guidance_loss.py:10-22 (TargetGuidance, per-sample map of the B=1 rule),
//           control/guidance.py:35-59 (autograd.grad wrt [x_guidance, action], scaled update, clip).
#include "common.cuh"

namespace b2p {

constexpr int TP_D = 64;      // hidden
constexpr int TP_H = 4;       // heads
constexpr int TP_HD = 16;     // head dim
constexpr int TP_FF = 256;
constexpr int TP_MAXS = 16;   // tokens (horizon - 1 <= 16)
constexpr int TP_NT = 512;
constexpr int TP_QS = 3 * TP_D + 4;   // row stride of the qkv buffers: 196 = 4 (mod 32), so the score loops (one key row per lane) spread over 8 banks instead of 1

// Shared-memory plan: 103.5 KB, so TWO trajectories (CTAs) share an SM and 256 of them fit the chip in one wave.  Only what
// the backward pass needs is kept per layer (qkv, attention probabilities, the normalised LayerNorm inputs and the FFN
// input x1 — the FFN pre-activation is recomputed from x1 in the backward pass, bit-identical to the forward value);
// everything else lives in scratch buffers that the forward and the backward pass use for different things.
struct LayerAct {            // saved for backward (floats, S = tokens)
  float qkv[TP_MAXS * TP_QS];
  float P[TP_H * TP_MAXS * TP_MAXS];
  float xh1[TP_MAXS * TP_D];   // normalised (pre-affine) LN1
  float x1[TP_MAXS * TP_D];    // LN1 output = FFN input
  float xh2[TP_MAXS * TP_D];   // normalised LN2
  float rstd1[TP_MAXS], rstd2[TP_MAXS];
};
struct TpSmem {
  LayerAct L[2];
  float xhf[TP_MAXS * TP_D];   // normalised final LayerNorm
  float rstdf[TP_MAXS];
  float out[TP_MAXS * 4];
  float act[TP_MAXS * 4];      // action rows (3 used)
  float te[TP_D];
  float ga[TP_MAXS * TP_FF];   // forward: attention output (first S*D), then FFN hidden.  backward: d(FFN hidden), then d(qkv)
  float gb[TP_MAXS * TP_FF];   // forward: final LayerNorm input / output.  backward: recomputed FFN pre-activation, then d(attention output)
  float gx[TP_MAXS * TP_D];    // forward: input of layer 0.  backward: running gradient wrt the layer input
  float gP[TP_H * TP_MAXS * TP_MAXS];   // forward: input of layer 1.  backward: d(P)
  int idx;
  __device__ __forceinline__ float* xin(int l) { return l == 0 ? gx : gP; }
  __device__ __forceinline__ float* att() { return ga; }
  __device__ __forceinline__ float* gqkv() { return ga; }
  __device__ __forceinline__ float* xfin() { return gb; }
  __device__ __forceinline__ float* xf() { return gb + TP_MAXS * TP_D; }
};
static_assert(sizeof(TpSmem) <= 110 * 1024, "two TrajPredict CTAs must fit one SM");

// Y[s][n] = sum_k X[s][k] * Wt[k][n] + b[n] on the tensor cores: the <= 16 token rows of a trajectory are exactly one m16 tile, so every
// linear layer of the forward AND the backward pass is a row of mma.sync.m16n8k8 TF32 tiles — in the 3xTF32 form (operands split into a
// TF32 high part and a TF32 residual, products hi*hi + lo*hi + hi*lo accumulated in fp32: fp32-class accuracy, the guidance gradient keeps
// its 1e-4 parity).  A warp owns N / 8 / 16 (rounded up) column tiles of 8 outputs.  The weights are read straight from L2 (the 400 KB do
// not fit beside the saved activations) as B fragments, up to 32 values per thread in flight before the first MMA (one L2 round trip per
// 64 - 128 values of K); the activations come from shared memory as float4 (K is walked in blocks of 16 with the k index permuted inside a
// block — lane t of a quad holds k0 + 4t .. 4t + 3 — so a row's fragment is one 16-byte load; A and B use the same permutation).
// Rows >= S hold whatever the buffer held: an output row depends on its own input row only, and rows >= S are never stored.
__device__ __forceinline__ uint32_t tf32_hi(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool ACCUM, int N>
__device__ __forceinline__ void linear(const float* __restrict__ X, int ldx, const float* __restrict__ Wt, const float* __restrict__ b,
                                       float* __restrict__ Y, int ldy, int S, int K) {
  constexpr int NT = N / 8;                        // column tiles
  constexpr int TPW = (NT + 15) / 16;              // tiles per warp: 1 (N = 64), 2 (N = 192, 256)
  constexpr int GB = TPW == 1 ? 8 : 4;             // K blocks of 16 whose weights are fetched together (<= 32 values per thread)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tile0 = warp * TPW;
  if (tile0 >= NT) return;
  const int ntl = NT - tile0 < TPW ? NT - tile0 : TPW;     // warp-uniform
  float acc[TPW][4];
#pragma unroll
  for (int i = 0; i < TPW; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  const float* xr0 = X + g * ldx + 4 * t;
  const float* xr1 = X + (g + 8) * ldx + 4 * t;
  const int nblk = K >> 4;
#pragma unroll 1
  for (int kb0 = 0; kb0 < nblk; kb0 += GB) {
    float w[TPW][GB][4];
#pragma unroll
    for (int i = 0; i < TPW; ++i)
#pragma unroll
      for (int kb = 0; kb < GB; ++kb)
        if (i < ntl && kb0 + kb < nblk) {
          const float* wp = Wt + (size_t)((kb0 + kb) * 16 + 4 * t) * N + (tile0 + i) * 8 + g;
#pragma unroll
          for (int j = 0; j < 4; ++j) w[i][kb][j] = __ldg(wp + (size_t)j * N);
        }
#pragma unroll
    for (int kb = 0; kb < GB; ++kb) {
      if (kb0 + kb < nblk) {
        const float4 x0 = *reinterpret_cast<const float4*>(xr0 + (kb0 + kb) * 16);
        const float4 x1 = *reinterpret_cast<const float4*>(xr1 + (kb0 + kb) * 16);
        const float xa[2][4] = {{x0.x, x1.x, x0.y, x1.y}, {x0.z, x1.z, x0.w, x1.w}};   // fragment order a0..a3 of the two k8 steps
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t ah[4], al[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) { ah[q] = tf32_hi(xa[h][q]); al[q] = tf32_hi(xa[h][q] - __uint_as_float(ah[q])); }
#pragma unroll
          for (int i = 0; i < TPW; ++i)
            if (i < ntl) {
              const float b0f = w[i][kb][2 * h], b1f = w[i][kb][2 * h + 1];
              const uint32_t b0h = tf32_hi(b0f), b1h = tf32_hi(b1f);
              const uint32_t b0l = tf32_hi(b0f - __uint_as_float(b0h)), b1l = tf32_hi(b1f - __uint_as_float(b1h));
              mma_tf32(acc[i], al, b0h, b1h);      // small terms first
              mma_tf32(acc[i], ah, b0l, b1l);
              mma_tf32(acc[i], ah, b0h, b1h);
            }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TPW; ++i)
    if (i < ntl) {
      const int n = (tile0 + i) * 8 + 2 * t;
      const float bb0 = b ? __ldg(b + n) : 0.f, bb1 = b ? __ldg(b + n + 1) : 0.f;
      if (g < S) {
        float* y = Y + g * ldy + n;
        if (ACCUM) { y[0] += acc[i][0] + bb0; y[1] += acc[i][1] + bb1; } else { y[0] = acc[i][0] + bb0; y[1] = acc[i][1] + bb1; }
      }
      if (g + 8 < S) {
        float* y = Y + (g + 8) * ldy + n;
        if (ACCUM) { y[0] += acc[i][2] + bb0; y[1] += acc[i][3] + bb1; } else { y[0] = acc[i][2] + bb0; y[1] = acc[i][3] + bb1; }
      }
    }
}

// y = LN(x) per row of 64; one warp per row.  Saves the normalised value and rstd when xh != null.
__device__ __forceinline__ void layer_norm(const float* x, const float* __restrict__ g, const float* __restrict__ b, float* y, float* xh,
                                           float* rstd_out, int S) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < S; s += TP_NT / 32) {
    float a0 = x[s * TP_D + lane], a1 = x[s * TP_D + 32 + lane];
    float sum = a0 + a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float mean = sum * (1.f / TP_D);
    float d0 = a0 - mean, d1 = a1 - mean;
    float sq = d0 * d0 + d1 * d1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    float rstd = 1.0f / sqrtf(sq * (1.f / TP_D) + 1e-5f);
    float h0 = d0 * rstd, h1 = d1 * rstd;
    if (xh) { xh[s * TP_D + lane] = h0; xh[s * TP_D + 32 + lane] = h1; if (lane == 0) rstd_out[s] = rstd; }
    y[s * TP_D + lane] = h0 * __ldg(g + lane) + __ldg(b + lane);
    y[s * TP_D + 32 + lane] = h1 * __ldg(g + 32 + lane) + __ldg(b + 32 + lane);
  }
}

// dx = rstd * (dyg - mean(dyg) - xh * mean(dyg*xh)),  dyg = dy * gamma.  In place on dy.
__device__ __forceinline__ void layer_norm_bwd(float* dy, const float* xh, const float* rstd, const float* __restrict__ g, int S) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < S; s += TP_NT / 32) {
    float g0 = dy[s * TP_D + lane] * __ldg(g + lane), g1 = dy[s * TP_D + 32 + lane] * __ldg(g + 32 + lane);
    float h0 = xh[s * TP_D + lane], h1 = xh[s * TP_D + 32 + lane];
    float m1 = g0 + g1, m2 = g0 * h0 + g1 * h1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
    m1 *= (1.f / TP_D); m2 *= (1.f / TP_D);
    float r = rstd[s];
    dy[s * TP_D + lane] = r * (g0 - m1 - h0 * m2);
    dy[s * TP_D + 32 + lane] = r * (g1 - m1 - h1 * m2);
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ void encoder_layer_fwd(TpSmem& sm, int l, const TrajPredWeights::Layer& w, float* xout, int S) {
  LayerAct& A = sm.L[l];
  const float* xin = sm.xin(l);
  float* att = sm.att();
  linear<false, 3 * TP_D>(xin, TP_D, w.qkv_wt, w.qkv_b, A.qkv, TP_QS, S, TP_D);
  __syncthreads();
  for (int i = threadIdx.x; i < TP_H * S * S; i += TP_NT) {   // scores
    int h = i / (S * S), r = (i / S) % S, c = i % S;
    const float* q = A.qkv + r * TP_QS + h * TP_HD;
    const float* k = A.qkv + c * TP_QS + TP_D + h * TP_HD;
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < TP_HD; ++e) s = fmaf(q[e], k[e], s);
    A.P[(h * TP_MAXS + r) * TP_MAXS + c] = s * 0.25f;         // 1/sqrt(16)
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TP_H * S; i += TP_NT) {       // softmax rows
    float* p = A.P + (size_t)(i / S * TP_MAXS + i % S) * TP_MAXS;
    float m = p[0];
    for (int c = 1; c < S; ++c) m = fmaxf(m, p[c]);
    float sum = 0.f;
    for (int c = 0; c < S; ++c) { float e = expf(p[c] - m); p[c] = e; sum += e; }
    float inv = 1.f / sum;
    for (int c = 0; c < S; ++c) p[c] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * TP_D; i += TP_NT) {       // att = P v
    int r = i / TP_D, col = i % TP_D, h = col / TP_HD;
    const float* p = A.P + (size_t)(h * TP_MAXS + r) * TP_MAXS;
    float s = 0.f;
    for (int c = 0; c < S; ++c) s = fmaf(p[c], A.qkv[c * TP_QS + 2 * TP_D + col], s);
    att[i] = s;
  }
  __syncthreads();
  // y1 = xin + att Wo^T + bo   (into x1 as scratch), then LN1
  linear<false, TP_D>(att, TP_D, w.out_wt, w.out_b, A.x1, TP_D, S, TP_D);
  __syncthreads();
  for (int i = threadIdx.x; i < S * TP_D; i += TP_NT) A.x1[i] += xin[i];
  __syncthreads();
  layer_norm(A.x1, w.n1_g, w.n1_b, A.x1, A.xh1, A.rstd1, S);
  __syncthreads();
  linear<false, TP_FF>(A.x1, TP_D, w.l1_wt, w.l1_b, sm.ga, TP_FF, S, TP_D);   // FFN pre-activation (the attention output in ga is dead)
  __syncthreads();
  for (int i = threadIdx.x; i < S * TP_FF; i += TP_NT) { float v = sm.ga[i]; sm.ga[i] = v * sigmoidf_(v); }  // SiLU, in place
  __syncthreads();
  linear<false, TP_D>(sm.ga, TP_FF, w.l2_wt, w.l2_b, xout, TP_D, S, TP_FF);
  __syncthreads();
  for (int i = threadIdx.x; i < S * TP_D; i += TP_NT) xout[i] += A.x1[i];
  __syncthreads();
  layer_norm(xout, w.n2_g, w.n2_b, xout, A.xh2, A.rstd2, S);
  __syncthreads();
}

// gradient wrt the layer output is in sm.gx on entry; gradient wrt the layer input is in sm.gx on exit
__device__ void encoder_layer_bwd(TpSmem& sm, int l, const TrajPredWeights::Layer& w, int S) {
  LayerAct& A = sm.L[l];
  float* gqkv = sm.gqkv();
  layer_norm_bwd(sm.gx, A.xh2, A.rstd2, w.n2_g, S);            // gx = d(x1 + ff)
  __syncthreads();
  linear<false, TP_FF>(sm.gx, TP_D, w.l2_w, nullptr, sm.ga, TP_FF, S, TP_D);   // d silu(h1) = gx W2   (W2 raw [64][256])
  linear<false, TP_FF>(A.x1, TP_D, w.l1_wt, w.l1_b, sm.gb, TP_FF, S, TP_D);    // h1 recomputed (same code, same inputs as the forward pass)
  __syncthreads();
  for (int i = threadIdx.x; i < S * TP_FF; i += TP_NT) {
    float v = sm.gb[i], sg = sigmoidf_(v);
    sm.ga[i] *= sg * (1.f + v * (1.f - sg));
  }
  __syncthreads();
  linear<true, TP_D>(sm.ga, TP_FF, w.l1_w, nullptr, sm.gx, TP_D, S, TP_FF);    // gx += d_h1 W1  (raw [256][64])  => d x1
  __syncthreads();
  layer_norm_bwd(sm.gx, A.xh1, A.rstd1, w.n1_g, S);            // gx = d(xin + o)
  __syncthreads();
  linear<false, TP_D>(sm.gx, TP_D, w.out_w, nullptr, sm.gb, TP_D, S, TP_D);    // d att = gx Wo  (raw [64][64])
  __syncthreads();
  for (int i = threadIdx.x; i < TP_H * S * S; i += TP_NT) {    // dP and dV
    int h = i / (S * S), r = (i / S) % S, c = i % S;
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < TP_HD; ++e) s = fmaf(sm.gb[r * TP_D + h * TP_HD + e], A.qkv[c * TP_QS + 2 * TP_D + h * TP_HD + e], s);
    sm.gP[(h * TP_MAXS + r) * TP_MAXS + c] = s;
  }
  for (int i = threadIdx.x; i < S * TP_D; i += TP_NT) {        // dV[c][col] = sum_r P[h][r][c] datt[r][col]
    int c = i / TP_D, col = i % TP_D, h = col / TP_HD;
    float s = 0.f;
    for (int r = 0; r < S; ++r) s = fmaf(A.P[(h * TP_MAXS + r) * TP_MAXS + c], sm.gb[r * TP_D + col], s);
    gqkv[c * TP_QS + 2 * TP_D + col] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TP_H * S; i += TP_NT) {        // softmax backward per row -> dS (scaled by 1/4)
    size_t o = (size_t)(i / S * TP_MAXS + i % S) * TP_MAXS;
    float dot = 0.f;
    for (int c = 0; c < S; ++c) dot = fmaf(A.P[o + c], sm.gP[o + c], dot);
    for (int c = 0; c < S; ++c) sm.gP[o + c] = A.P[o + c] * (sm.gP[o + c] - dot) * 0.25f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * TP_D; i += TP_NT) {        // dQ and dK
    int r = i / TP_D, col = i % TP_D, h = col / TP_HD;
    float dq = 0.f, dk = 0.f;
    for (int c = 0; c < S; ++c) {
      dq = fmaf(sm.gP[(h * TP_MAXS + r) * TP_MAXS + c], A.qkv[c * TP_QS + TP_D + col], dq);
      dk = fmaf(sm.gP[(h * TP_MAXS + c) * TP_MAXS + r], A.qkv[c * TP_QS + col], dk);
    }
    gqkv[r * TP_QS + col] = dq;
    gqkv[r * TP_QS + TP_D + col] = dk;
  }
  __syncthreads();
  linear<true, TP_D>(gqkv, TP_QS, w.qkv_w, nullptr, sm.gx, TP_D, S, 3 * TP_D);  // gx += dqkv Wqkv (raw [192][64])
  __syncthreads();
}

// action: [B, H, action_stride] rows, 3 used columns at action_col0.
// GUIDE == false: writes state (dense [B,S,SD] if !full, else full model_output [B,H,D] with zero row + action).
// GUIDE == true : `out` is the model_output [B,H,D]; if state_given the state columns already in `out` are used for the
//                 index rule / state gradient, else they are written first.  Then the guidance update is applied in place.
// MODE 0: forward only; 1: guidance update; 2: VJP with an arbitrary cotangent `target` = grad_state [B,S,SD] -> out = grad_action [B,H,3]
template <int MODE>
__global__ void __launch_bounds__(TP_NT, 2) trajpred_kernel(TrajPredWeights w, const float* __restrict__ action, int action_stride,
                                                         int action_col0, const float* __restrict__ time_embed, int te_stride, float* out, int full,
                                                         int state_given, const float* __restrict__ target, float grad_scale,
                                                         float scale_state, float scale_action, int H, int D) {
  extern __shared__ __align__(16) unsigned char raw[];
  TpSmem& sm = *reinterpret_cast<TpSmem*>(raw);
  const int b = blockIdx.x, tid = threadIdx.x, S = H - 1, SD = D - 3;
  for (int i = tid; i < S * 3; i += TP_NT) sm.act[(i / 3) * 4 + i % 3] = action[((size_t)b * H + i / 3) * action_stride + action_col0 + i % 3];
  if (tid < TP_D) sm.te[tid] = time_embed[(size_t)b * te_stride + tid];
  __syncthreads();
  for (int i = tid; i < S * TP_D; i += TP_NT) {   // x0 = input_proj(action) + pos + time_embed
    int s = i / TP_D, n = i % TP_D;
    float v = __ldg(w.in_b + n);
    v = fmaf(sm.act[s * 4 + 0], __ldg(w.in_w + 0 * TP_D + n), v);
    v = fmaf(sm.act[s * 4 + 1], __ldg(w.in_w + 1 * TP_D + n), v);
    v = fmaf(sm.act[s * 4 + 2], __ldg(w.in_w + 2 * TP_D + n), v);
    sm.xin(0)[i] = v + __ldg(w.pos + i) + sm.te[n];
  }
  __syncthreads();
  encoder_layer_fwd(sm, 0, w.layer[0], sm.xin(1), S);
  encoder_layer_fwd(sm, 1, w.layer[1], sm.xfin(), S);
  layer_norm(sm.xfin(), w.fn_g, w.fn_b, sm.xf(), sm.xhf, sm.rstdf, S);
  __syncthreads();
  const float* xf = sm.xf();
  for (int i = tid; i < S * SD; i += TP_NT) {
    int s = i / SD, c = i % SD;
    float v = __ldg(w.out_b + c);
    for (int k = 0; k < TP_D; ++k) v = fmaf(xf[s * TP_D + k], __ldg(w.out_wt + k * SD + c), v);
    sm.out[s * 4 + c] = v;
  }
  __syncthreads();
  if (MODE == 0) {
    if (!full) {
      for (int i = tid; i < S * SD; i += TP_NT) out[(size_t)b * S * SD + i] = sm.out[(i / SD) * 4 + i % SD];
    } else {
      for (int i = tid; i < H * D; i += TP_NT) {
        int r = i / D, c = i % D;
        float v;
        if (c < SD) v = r == 0 ? 0.f : sm.out[(r - 1) * 4 + c];
        else v = action[((size_t)b * H + r) * action_stride + action_col0 + (c - SD)];
        out[(size_t)b * H * D + i] = v;
      }
    }
    return;
  }
  if (MODE == 2) {   // generic VJP: gx = cot @ W_out, backward, d_action = gx @ W_in
    const float* cot = target + (size_t)b * S * SD;
    for (int i = tid; i < S * TP_D; i += TP_NT) {
      int s = i / TP_D, k = i % TP_D;
      float v = 0.f;
      for (int c = 0; c < SD; ++c) v = fmaf(cot[s * SD + c], __ldg(w.out_w + c * TP_D + k), v);
      sm.gx[i] = v;
    }
    __syncthreads();
    layer_norm_bwd(sm.gx, sm.xhf, sm.rstdf, w.fn_g, S);
    __syncthreads();
    encoder_layer_bwd(sm, 1, w.layer[1], S);
    encoder_layer_bwd(sm, 0, w.layer[0], S);
    for (int i = tid; i < H * 3; i += TP_NT) {
      int r = i / 3, a = i % 3;
      float g = 0.f;
      if (r < S) for (int n = 0; n < TP_D; ++n) g = fmaf(sm.gx[r * TP_D + n], __ldg(w.in_w_raw + n * 3 + a), g);
      out[(size_t)b * H * 3 + i] = g;
    }
    return;
  }
  // ------------------------------- guidance -------------------------------
  float* mo = out + (size_t)b * H * D;
  if (!state_given) {
    for (int i = tid; i < H * D; i += TP_NT) {
      int r = i / D, c = i % D;
      float v;
      if (c < SD) v = r == 0 ? 0.f : sm.out[(r - 1) * 4 + c];
      else v = action[((size_t)b * H + r) * action_stride + action_col0 + (c - SD)];
      mo[i] = v;
    }
    __syncthreads();
  }
  const float tx = target[b * 2 + 0], ty = target[b * 2 + 1];
  if (tid == 0) {   // TargetGuidance index rule (guidance_loss.py:16-21)
    float x0 = mo[0], y0 = mo[1];
    float t2a = sqrtf((tx - x0) * (tx - x0) + (ty - y0) * (ty - y0));
    float fx = mo[(H - 1) * D] - x0, fy = mo[(H - 1) * D + 1] - y0;
    float f2a = sqrtf(fx * fx + fy * fy);
    int idx = 0;
    if (!(f2a < t2a)) {
      float best = INFINITY;
      for (int r = 0; r < H; ++r) {
        float dx = mo[r * D] - tx, dy = mo[r * D + 1] - ty;
        float d = dx * dx + dy * dy;
        if (d < best) { best = d; idx = r; }
      }
    }
    sm.idx = idx;
  }
  __syncthreads();
  const int idx = sm.idx;
  const float gsx = 2.f * (mo[idx * D + 0] - tx), gsy = 2.f * (mo[idx * D + 1] - ty);   // d loss / d x[idx, :2]
  // ---- VJP through TrajPredict: cotangent only at output row idx-1, cols 0..1 ----
  for (int i = tid; i < S * TP_D; i += TP_NT) {
    int s = i / TP_D, k = i % TP_D;
    float v = 0.f;
    if (idx >= 1 && s == idx - 1) v = gsx * __ldg(w.out_w + 0 * TP_D + k) + gsy * __ldg(w.out_w + 1 * TP_D + k);
    sm.gx[i] = v;
  }
  __syncthreads();
  if (idx >= 1) {   // block-uniform branch
    layer_norm_bwd(sm.gx, sm.xhf, sm.rstdf, w.fn_g, S);
    __syncthreads();
    encoder_layer_bwd(sm, 1, w.layer[1], S);
    encoder_layer_bwd(sm, 0, w.layer[0], S);
  }
  // ---- update (control/guidance.py:52-59): state cols use scale/15, action cols use scale; then clip ----
  for (int i = tid; i < H * D; i += TP_NT) {
    int r = i / D, c = i % D;
    float g = 0.f;
    if (c < SD) {
      if (r == idx && c == 0) g = gsx;
      if (r == idx && c == 1) g = gsy;
    } else if (r < S && idx >= 1) {
      int a = c - SD;   // d action[r][a] = sum_n gx[r][n] * in_w_raw[n][a]
      for (int n = 0; n < TP_D; ++n) g = fmaf(sm.gx[r * TP_D + n], __ldg(w.in_w_raw + n * 3 + a), g);
    }
    g *= grad_scale;
    float v = mo[i] - (c < SD ? scale_state : scale_action) * g;
    mo[i] = fminf(fmaxf(v, -1.f), 1.f);
  }
}

static int tp_check(int H, int D) { return (H - 1 > TP_MAXS || H < 2 || D - 3 > 4 || D - 3 < 2) ? B2P_ERR_INVALID_ARG : B2P_OK; }

int launch_state_pred(const TrajPredWeights& w, const float* action, const float* time_embed, int te_stride, float* out, int full_output, int B,
                      int H, int D, cudaStream_t s) {
  if (tp_check(H, D)) return B2P_ERR_INVALID_ARG;
  B2P_CUDA_TRY(cudaFuncSetAttribute(trajpred_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TpSmem)));
  trajpred_kernel<0><<<B, TP_NT, sizeof(TpSmem), s>>>(w, action, 3, 0, time_embed, te_stride, out, full_output, 0, nullptr, 0.f, 0.f, 0.f, H, D);
  return (int)cudaGetLastError();
}

int launch_classifier_guidance(const TrajPredWeights& w, float* model_output, const float* time_embed, int te_stride, const float* target,
                               float grad_scale, float scale, int B, int H, int D, cudaStream_t s) {
  if (tp_check(H, D)) return B2P_ERR_INVALID_ARG;
  B2P_CUDA_TRY(cudaFuncSetAttribute(trajpred_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TpSmem)));
  float scale_state = (float)((double)scale / 15.0);   // quirk 5 (control/guidance.py:56)
  trajpred_kernel<1><<<B, TP_NT, sizeof(TpSmem), s>>>(w, model_output, D, D - 3, time_embed, te_stride, model_output, 1, 1, target, grad_scale,
                                                         scale_state, scale, H, D);
  return (int)cudaGetLastError();
}

int launch_classifier_guidance_from_action(const TrajPredWeights& w, const float* action, float* model_output, const float* time_embed,
                                           int te_stride, const float* target, float grad_scale, float scale, int B, int H, int D, cudaStream_t s) {
  if (tp_check(H, D)) return B2P_ERR_INVALID_ARG;
  B2P_CUDA_TRY(cudaFuncSetAttribute(trajpred_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TpSmem)));
  float scale_state = (float)((double)scale / 15.0);   // quirk 5 (control/guidance.py:56)
  trajpred_kernel<1><<<B, TP_NT, sizeof(TpSmem), s>>>(w, action, 3, 0, time_embed, te_stride, model_output, 1, 0, target, grad_scale,
                                                         scale_state, scale, H, D);
  return (int)cudaGetLastError();
}

int launch_state_pred_vjp(const TrajPredWeights& w, const float* action, const float* time_embed, int te_stride, const float* grad_state,
                          float* grad_action, int B, int H, int D, cudaStream_t s) {
  if (tp_check(H, D)) return B2P_ERR_INVALID_ARG;
  B2P_CUDA_TRY(cudaFuncSetAttribute(trajpred_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TpSmem)));
  trajpred_kernel<2><<<B, TP_NT, sizeof(TpSmem), s>>>(w, action, 3, 0, time_embed, te_stride, grad_action, 0, 0, grad_state, 0.f, 0.f, 0.f, H, D);
  return (int)cudaGetLastError();
}

}  // namespace b2p
