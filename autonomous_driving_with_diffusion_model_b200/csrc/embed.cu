// Time / condition embedding, one CTA per row.  Replaces
//   SinusoidalPosEmb -> Linear -> Mish -> Linear           (modeling/temporal.py:93-98, helpers.py:67-74)
//   time_embed += cond_mlp(cond)  (FREE_GUIDANCE)           (modeling/temporal.py:206-212)
//   cat[time_embed, img_feature] and the Mish that heads every block's time_mlp (temporal.py:213, 35-39)
// Outputs the two halves of Mish(cat[time_embed, feat]) as separate dense matrices, because
//   Linear(Mish(cat[te, f])) == W[:, :dim] Mish(te) + W[:, dim:] Mish(f) + b            
// lets the 16 per-block Linear(2*dim -> C_out) run as ONE GEMM over the concatenated weight — either on both halves
// (general forward) or, inside a plan, as a [T, sum C_out] time table built once plus a [B, sum C_out] image term
// built once (the feature is step-invariant).
#include "common.cuh"

namespace b2p {

__constant__ float c_freq[512];

int upload_freq_table(const float* f, int n) {
  return (int)cudaMemcpyToSymbol(c_freq, f, sizeof(float) * n);
}

// blockDim.x == 4*dim.  Dot products are split 4-way over the threads and combined through shared memory so that every
// thread has dim/… independent loads in flight (the naive one-thread-per-output loop is L2-latency bound).
__global__ void embed_kernel(EmbedArgs a) {
  extern __shared__ float sh[];
  const int dim = a.dim, dim4 = 4 * a.dim, half = a.dim / 2;
  float* e0 = sh;              // [dim]
  float* h = e0 + dim;         // [4dim]
  float* c1 = h + dim4;        // [dim]
  float* part = c1 + dim;      // [4][dim] partial sums
  float* part2 = part + dim4;  // [4][dim]
  const int b = blockIdx.x, tid = threadIdx.x;
  const bool do_te = b < a.te_rows;
  pdl_wait();
  pdl_launch_dependents();
  if (do_te) {
    const float t = (float)a.t[b % a.t_count];
    if (tid < dim) {
      float arg = t * c_freq[tid % half];
      e0[tid] = tid < half ? sinf(arg) : cosf(arg);
    }
    if (a.use_cond && tid >= dim && tid < 2 * dim) {   // cond_mlp.0 (2 -> dim) + Mish
      const int o = tid - dim;
      float s = __ldg(a.bc0 + o);
      if (a.cond) {
        s = fmaf(__ldg(a.wc0t + o), a.cond[b * 2 + 0], s);
        s = fmaf(__ldg(a.wc0t + dim + o), a.cond[b * 2 + 1], s);
      }
      c1[o] = mish_f(s);
    }
    __syncthreads();
    {   // h = Mish(W1 e0 + b1): one output per thread, dim-long dot, loads fully unrolled
      float s = __ldg(a.b1 + tid);
#pragma unroll 16
      for (int i = 0; i < dim; ++i) s = fmaf(__ldg(a.w1t + i * dim4 + tid), e0[i], s);
      h[tid] = mish_f(s);
    }
    __syncthreads();
    {   // te = W3 h + b3 (+ Wc2 c1 + bc2): output o = tid % dim, quarter q = tid / dim of the reduction
      const int o = tid % dim, q = tid / dim;
      float s = 0.f;
#pragma unroll 16
      for (int i = q * dim; i < (q + 1) * dim; ++i) s = fmaf(__ldg(a.w3t + i * dim + o), h[i], s);
      part[q * dim + o] = s;
      if (a.use_cond) {
        float c = 0.f;
        const int n = dim / 4;
#pragma unroll 16
        for (int i = q * n; i < (q + 1) * n; ++i) c = fmaf(__ldg(a.wc2t + i * dim + o), c1[i], c);
        part2[q * dim + o] = c;
      }
    }
    __syncthreads();
    if (tid < dim) {
      float s = __ldg(a.b3 + tid) + ((part[tid] + part[dim + tid]) + (part[2 * dim + tid] + part[3 * dim + tid]));
      if (a.use_cond) s += __ldg(a.bc2 + tid) + ((part2[tid] + part2[dim + tid]) + (part2[2 * dim + tid] + part2[3 * dim + tid]));
      if (a.time_embed) a.time_embed[(size_t)b * dim + tid] = s;
      a.mish_te[(size_t)b * dim + tid] = mish_f(s);
    }
  }
  if (b < a.feat_out_rows && tid >= dim && tid < 2 * dim) {
    const int o = tid - dim;
    a.mish_feat[(size_t)b * dim + o] = mish_f(__ldg(a.feat + (size_t)(b % a.feat_rows) * dim + o));
  }
}

int launch_embed(const EmbedArgs& a, cudaStream_t s) {
  if (a.dim * 4 > 1024 || a.dim % 4 != 0 || a.B <= 0 || a.t_count <= 0 || a.feat_rows <= 0) return B2P_ERR_INVALID_ARG;
  prefer_max_smem_carveout((const void*)embed_kernel);
  embed_kernel<<<a.B, 4 * a.dim, sizeof(float) * 14 * a.dim, s>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace b2p
