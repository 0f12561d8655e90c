// Time / condition embedding, one CTA per row of the denoiser batch.  Replaces
//   SinusoidalPosEmb -> Linear -> Mish -> Linear           (modeling/temporal.py:93-98, helpers.py:67-74)
//   time_embed += cond_mlp(cond)  (FREE_GUIDANCE)           (modeling/temporal.py:206-212)
//   cat[time_embed, img_feature] and the Mish that heads every block's time_mlp (temporal.py:213, 35-39)
// The 16 per-block Linear(128 -> C_out) are then ONE GEMM over the concatenated [128][sum C_out] weight
// (launched through the conv kernel with L = 1).
#include "common.cuh"

namespace b2p {

__constant__ float c_freq[512];

int upload_freq_table(const float* f, int n) {
  return (int)cudaMemcpyToSymbol(c_freq, f, sizeof(float) * n);
}

__global__ void embed_kernel(EmbedArgs a) {
  extern __shared__ float sh[];
  const int dim = a.dim, dim4 = 4 * a.dim, half = a.dim / 2;
  float* e0 = sh;            // [dim]
  float* h = e0 + dim;       // [4dim]
  float* te = h + dim4;      // [dim]
  float* c1 = te + dim;      // [dim]
  const int b = blockIdx.x, tid = threadIdx.x;
  const float t = (float)a.t[b % a.t_count];
  if (tid < dim) {
    float arg = t * c_freq[tid % half];
    e0[tid] = tid < half ? sinf(arg) : cosf(arg);
  }
  __syncthreads();
  if (tid < dim4) {
    float s = __ldg(a.b1 + tid);
    for (int i = 0; i < dim; ++i) s = fmaf(__ldg(a.w1t + i * dim4 + tid), e0[i], s);
    h[tid] = mish_f(s);
  }
  if (a.use_cond && tid < dim) {
    float s = __ldg(a.bc0 + tid);
    if (a.cond) {
      s = fmaf(__ldg(a.wc0t + tid), a.cond[b * 2 + 0], s);
      s = fmaf(__ldg(a.wc0t + dim + tid), a.cond[b * 2 + 1], s);
    }
    c1[tid] = mish_f(s);
  }
  __syncthreads();
  if (tid < dim) {
    float s = __ldg(a.b3 + tid);
    for (int i = 0; i < dim4; ++i) s = fmaf(__ldg(a.w3t + i * dim + tid), h[i], s);
    if (a.use_cond) {
      float c = __ldg(a.bc2 + tid);
      for (int i = 0; i < dim; ++i) c = fmaf(__ldg(a.wc2t + i * dim + tid), c1[i], c);
      s += c;
    }
    if (a.time_embed) a.time_embed[(size_t)b * dim + tid] = s;
    a.mish_cond[(size_t)b * 2 * dim + tid] = mish_f(s);
    a.mish_cond[(size_t)b * 2 * dim + dim + tid] = mish_f(__ldg(a.feat + (size_t)(b % a.feat_rows) * dim + tid));
  }
}

int launch_embed(const EmbedArgs& a, cudaStream_t s) {
  if (a.dim * 4 > 1024 || a.dim > 512 || a.B <= 0 || a.t_count <= 0 || a.feat_rows <= 0) return B2P_ERR_INVALID_ARG;
  embed_kernel<<<a.B, 4 * a.dim, sizeof(float) * 7 * a.dim, s>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace b2p
