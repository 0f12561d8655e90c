// Internal declarations shared by the .cu files of libb200plan.  Not part of the public ABI (include/b200plan.h).
#pragma once
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/b200plan.h"

#define B2P_CUDA_TRY(expr)                    \
  do {                                        \
    cudaError_t _e = (expr);                  \
    if (_e != cudaSuccess) return (int)_e;    \
  } while (0)

namespace b2p {

// ---------------------------------------------------------------------------------------------------------
// Fused conv layer on channels-last activations [rows = (sample, position), channels].
// out[b, l, co] = epilogue( sum_{j, c} W[j][c][co] * in[b, pos(l, j), c] + bias[co] )
//   plain conv :   pos = l*stride + j - pad                (Conv1d,  modeling/helpers.py:77-83, 95-112)
//   transposed :   pos = (l + pad - j)/stride if divisible (ConvTranspose1d k4 s2 p1, modeling/helpers.py:86-92)
// The input may be the channel concatenation of two tensors (skip connection, modeling/temporal.py:227).
// Epilogue (all optional): GroupNorm(8)+Mish, + temb[b, co], + residual (identity or 1x1 conv of `res` inputs),
// fused 1x1 head (final_conv.1 / act_conv.1) written as [rows, head_dim].
// ---------------------------------------------------------------------------------------------------------
struct ConvArgs {
  const float* x0; const float* x1;   // inputs; x1 may be null
  int C0, C1;                         // channels of x0 / x1
  int x0_period, rx0_period;          // >0: sample b reads x0 row (b % period)  (CFG feeds [x; x] without a copy)
  int Lin, Lout, log2Lout;            // positions per sample
  int nrows;                          // B * Lout
  int Cout;
  int taps, jmin, jmax;               // kernel taps; only taps in [jmin, jmax] can touch a valid position
  int stride, pad, transposed;
  const float* W;                     // packed [taps][C0+C1][Cout]
  const float* Wk;                    // K-major copy [Cout][taps][C0+C1] (small-batch GEMV path), may be null
  const float* resWk;                 // K-major residual weights [Cout][RC0+RC1], may be null
  const float* bias;                  // [Cout]
  const float* gn_gamma; const float* gn_beta;  // null => no GroupNorm/Mish
  int cg;                             // channels per group (Cout/8)
  const float* temb; int temb_stride; // + temb[b*temb_stride + co]
  const float* temb2;                 // + temb2[co] (per-step time term shared by the whole batch), may be null
  const float* res_id;                // identity residual [nrows, Cout]
  const float* rx0; const float* rx1; int RC0, RC1;  // residual 1x1 conv inputs (same Lout rows)
  const float* resW; const float* resB;              // [RC0+RC1][Cout], [Cout]
  const float* headW; const float* headB; int head_dim; float* head_out;  // [64][head_dim]
  float* out;                         // [nrows, Cout] (may be null when only the head is wanted)
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;  // optional bf16 hi/lo copy of the output (tensor-core precisions)
  float* res_out;                     // if set, the residual 1x1 result is written here instead of being added
};

int launch_conv_ffma(const ConvArgs& a, cudaStream_t s);
// small-batch (<= 32 output rows) exact-fp32 GEMV path (conv_gemv.cu)
bool conv_gemv_applicable(const ConvArgs& a);
int launch_conv_gemv(const ConvArgs& a, cudaStream_t s);

// Ask for the maximum shared-memory carve-out for `kernel` (once per kernel).  Every kernel of the plan uses the same
// L1/shared split, so kernels that overlap under programmatic dependent launch never wait for an SM to be reconfigured.
inline void prefer_max_smem_carveout(const void* kernel) {
  static const void* seen[256];          // (kernel, device) pairs already configured
  static int seen_dev[256];
  static int nseen = 0;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("B2P_CARVEOUT"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  for (int i = 0; i < nseen; ++i) if (seen[i] == kernel && seen_dev[i] == dev) return;
  if (nseen < 256) { seen[nseen] = kernel; seen_dev[nseen] = dev; ++nseen; }
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// L2 residency policy of the packed weights (north_star: "L2-resident weights"): the tensor-core launches carry an access-policy
// window over the 16-bit weight pack (hit ratio 1, persisting; everything else streams), so activations of a large batch cannot
// evict the weights between the T iterations.  Set per thread by the API layer before it enqueues an evaluation; bytes == 0: none.
struct L2Window { const void* base; size_t bytes; };
inline L2Window& l2_weight_window() { static thread_local L2Window w{nullptr, 0}; return w; }
inline int add_l2_window_attr(cudaLaunchAttribute* attr, int n) {
  const L2Window& w = l2_weight_window();
  if (!w.bytes) return n;
  attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
  attr[n].val.accessPolicyWindow.base_ptr = const_cast<void*>(w.base);
  attr[n].val.accessPolicyWindow.num_bytes = w.bytes;
  static float ratio = -1.f;
  if (ratio < 0.f) { const char* e = getenv("B2P_L2_HITRATIO"); ratio = e ? (float)atof(e) : 1.0f; }
  attr[n].val.accessPolicyWindow.hitRatio = ratio;
  attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  return n + 1;
}

// launch `kernel` with the programmatic-stream-serialization attribute (the kernel must call pdl_wait() before it reads
// anything produced by the preceding kernel)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  prefer_max_smem_carveout((const void*)kernel);
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// time / condition embedding (modeling/temporal.py:205-213 + the Mish in front of every block's time_mlp)
struct EmbedArgs {
  const int64_t* t; int t_count;      // timesteps, repeated to B
  const float* feat; int feat_rows;   // [feat_rows, dim], repeated to B
  const float* cond;                  // [B,2] or null (zeros)
  int use_cond;                       // FREE_GUIDANCE
  const float* w1t; const float* b1;  // time_mlp.1 transposed [dim][4dim]
  const float* w3t; const float* b3;  // time_mlp.3 transposed [4dim][dim]
  const float* wc0t; const float* bc0; const float* wc2t; const float* bc2;  // cond_mlp transposed
  float* time_embed;                  // [te_rows, dim] (may be null)
  float* mish_te;                     // [te_rows, dim]   = Mish(time_embed)         } the two halves of
  float* mish_feat;                   // [feat_out_rows, dim] = Mish(feat)           } Mish(cat[time_embed, feat])
  int te_rows, feat_out_rows;         // rows written to each output (<= B)
  int B, dim;
};
int launch_embed(const EmbedArgs& a, cudaStream_t s);

// scheduler
struct SchedLaunch {
  b2p_sched_config sc; b2p_step_coeffs k;
  const float* mo; const float* mo_u; float cfg_scale;
  const float* sample; const float* noise; const float* traj; const float* mask;
  float* prev; float* x0; int B, H, D; float eta, magic; int flags;
  // in-kernel noise (K8 in SURVEY.md Appendix C): when `noise` is null and the step consumes noise, every thread draws its four
  // N(0,1) values from Philox4x32-10 keyed by *seed (read at run time: a captured graph sees each plan's seed) with the
  // counter (element group, noise_step)
  const unsigned long long* seed; unsigned noise_step;
  float* thr_scratch;   // [B] floats for the dynamic-threshold quantile (sample_max_value > 1); null => stream-ordered allocation
};
int launch_sched_step(const SchedLaunch& a, cudaStream_t s);

// TrajPredict forward / classifier guidance (trajpred.cu)
struct TrajPredWeights {
  const float* in_w; const float* in_b;          // [3][64] transposed, [64]
  const float* pos;                              // [S][64] sinusoidal table
  struct Layer {
    const float* qkv_wt; const float* qkv_b;     // [64][192], [192]
    const float* out_wt; const float* out_b;     // [64][64]
    const float* l1_wt; const float* l1_b;       // [64][256]
    const float* l2_wt; const float* l2_b;       // [256][64]
    const float* n1_g; const float* n1_b; const float* n2_g; const float* n2_b;
    // un-transposed copies for the backward pass
    const float* qkv_w; const float* out_w; const float* l1_w; const float* l2_w;
  } layer[4];
  int n_layers;
  const float* fn_g; const float* fn_b;          // final LayerNorm
  const float* out_wt; const float* out_b;       // [64][4] transposed
  const float* out_w;                            // [4][64]
  const float* in_w_raw;                         // [64][3]
};
// full_output == 0: out = state [B, H-1, D-3];  == 1: out = cat[cat[0, state], action] [B, H, D] (temporal.py:237-241)
// te_stride: row stride of time_embed in floats (dim, or 0 when one row is shared by the whole batch)
int launch_state_pred(const TrajPredWeights& w, const float* action, const float* time_embed, int te_stride, float* out, int full_output,
                      int B, int H, int D, cudaStream_t s);
int launch_state_pred_vjp(const TrajPredWeights& w, const float* action, const float* time_embed, int te_stride, const float* grad_state,
                          float* grad_action, int B, int H, int D, cudaStream_t s);
int launch_classifier_guidance(const TrajPredWeights& w, float* model_output, const float* time_embed, int te_stride, const float* target,
                               float grad_scale, float scale, int B, int H, int D, cudaStream_t s);
// same, but the model output is first assembled from the denoiser's action head: model_output = cat[cat[0, state_pred(action)], action]
// (modeling/temporal.py:237-241) — the TrajPredict forward runs once for both the output and the guidance gradient
int launch_classifier_guidance_from_action(const TrajPredWeights& w, const float* action, float* model_output, const float* time_embed,
                                           int te_stride, const float* target, float grad_scale, float scale, int B, int H, int D, cudaStream_t s);

// programmatic dependent launch (PDL): wait for the preceding kernel's results / let the following kernel start its prologue
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float mish_f(float x) {
  // x * tanh(softplus(x)), softplus threshold 20 (nn.Mish); tanh(log1p(e^x)) == n/(n+2), n = e^x (e^x + 2)
  if (x > 20.f) return x;
  float e = expf(x);
  float n = e * (e + 2.f);
  return x * (n / (n + 2.f));
}

}  // namespace b2p
