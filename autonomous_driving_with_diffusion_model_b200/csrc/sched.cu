// Fused scheduler step: CFG mix + x0/eps + clamp/threshold + DDIM/DDPM update + noise + inpainting blend +
// first-waypoint overwrite + final post-process, one elementwise launch (K1 in SURVEY.md Appendix C).
// Replaces scheduler/guidance_ddim_scheduler.py:60-173, guidance_ddpm_scheduler.py:59-178,
// inpainting_ddim_scheduler.py:10-153, inpainting_ddpm_scheduler.py:10-146, interact.py:142-144,164,166-167.
// Arithmetic uses the round-to-nearest intrinsics in the reference's operation order (no FMA contraction), so the
// result is bit-identical to the fp32 CPU oracle.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "sched_math.cuh"

namespace b2p {

__global__ void __launch_bounds__(256) sched_step_kernel(SchedK a) {
  pdl_wait();                 // model_output comes from the preceding kernel
  pdl_launch_dependents();
  int i4 = blockIdx.x * blockDim.x + threadIdx.x;
  int base = i4 * 4;
  if (base >= a.n) return;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 m = __ldg(reinterpret_cast<const float4*>(a.mo) + i4);
  float4 mu = a.mo_u ? __ldg(reinterpret_cast<const float4*>(a.mo_u) + i4) : z;
  float4 x = __ldg(reinterpret_cast<const float4*>(a.sample) + i4);
  float4 nz = a.noise ? __ldg(reinterpret_cast<const float4*>(a.noise) + i4) : z;
  if (!a.noise && a.seed) nz = philox_normal4(*a.seed, (unsigned)i4, a.noise_step);
  float4 tj = a.traj ? __ldg(reinterpret_cast<const float4*>(a.traj) + i4) : z;
  float4 mk = a.mask ? __ldg(reinterpret_cast<const float4*>(a.mask) + i4) : z;
  float4 o, x0;
  // H*D is a multiple of 4 (checked by the launcher), so the 4 elements of a thread belong to one trajectory
  const int sample = base / a.HD, pos = base - sample * a.HD;
  int c0 = pos % a.D, c1 = c0 + 1 == a.D ? 0 : c0 + 1, c2 = c1 + 1 == a.D ? 0 : c1 + 1, c3 = c2 + 1 == a.D ? 0 : c2 + 1;
  o.x = step_one(a, sample, pos + 0, c0, m.x, mu.x, x.x, nz.x, tj.x, mk.x, &x0.x);
  o.y = step_one(a, sample, pos + 1, c1, m.y, mu.y, x.y, nz.y, tj.y, mk.y, &x0.y);
  o.z = step_one(a, sample, pos + 2, c2, m.z, mu.z, x.z, nz.z, tj.z, mk.z, &x0.z);
  o.w = step_one(a, sample, pos + 3, c3, m.w, mu.w, x.w, nz.w, tj.w, mk.w, &x0.w);
  reinterpret_cast<float4*>(a.prev)[i4] = o;
  if (a.x0_out) reinterpret_cast<float4*>(a.x0_out)[i4] = x0;
}

// Dynamic thresholding (sample_max_value > 1): per-sample quantile of |x0| with torch.quantile's linear
// interpolation (guidance_ddim_scheduler.py:23-58; quirk 8: over all H*D elements).  One CTA per sample,
// rank-by-counting (n = H*D = 112 is tiny).
__global__ void __launch_bounds__(128) threshold_s_kernel(SchedK a, float ratio, float max_value, float* s_out) {
  extern __shared__ float v[];
  int b = blockIdx.x, n = a.HD;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int idx = b * n + i;
    float m = a.mo[idx];
    if (a.mo_u) { float mu = a.mo_u[idx]; m = add(mu, mul(a.cfg_scale, sub(m, mu))); }
    v[i] = fabsf(x0_of(a, m, a.sample[idx]));
  }
  __syncthreads();
  float rank = mul(ratio, (float)(n - 1));
  int lo = (int)floorf(rank);
  int hi = min(lo + 1, n - 1);
  float w = sub(rank, (float)lo);
  __shared__ float sel[2];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float vi = v[i];
    int r = 0;
    for (int j = 0; j < n; ++j) { float vj = v[j]; r += (vj < vi) || (vj == vi && j < i); }
    if (r == lo) sel[0] = vi;
    if (r == hi) sel[1] = vi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lo_v = sel[0], hi_v = sel[1];
    float d = sub(hi_v, lo_v);
    // at::lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
    float q = (w < 0.5f) ? add(lo_v, mul(w, d)) : sub(hi_v, mul(d, sub(1.f, w)));
    s_out[b] = fminf(fmaxf(q, 1.f), max_value);
  }
}

// the noise tensor [steps, n] a plan with noise == NULL consumes (same counters as sched_step_kernel: group = element / 4, step)
__global__ void __launch_bounds__(256) philox_fill_kernel(unsigned long long seed, int groups_per_step, int steps, float* out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x, st = blockIdx.y;
  if (g >= groups_per_step || st >= steps) return;
  reinterpret_cast<float4*>(out)[(size_t)st * groups_per_step + g] = philox_normal4(seed, (unsigned)g, (unsigned)st);
}

// kernel arguments of one step (shared with the seam kernel of chain64.cu, which runs the same arithmetic behind the denoiser's head)
int sched_make_args(const SchedLaunch& L, SchedK* out) {
  if (!L.sample || L.B <= 0) return B2P_ERR_INVALID_ARG;
  if ((L.H * L.D) % 4 != 0) return B2P_ERR_INVALID_ARG;
  SchedK a;
  memset(&a, 0, sizeof(a));
  a.mo = L.mo; a.mo_u = L.mo_u; a.cfg_scale = L.cfg_scale; a.sample = L.sample; a.noise = L.noise;
  a.traj = L.traj; a.mask = L.mask; a.prev = L.prev; a.x0_out = L.x0; a.n = L.B * L.H * L.D; a.HD = L.H * L.D; a.D = L.D;
  a.ddpm = (L.sc.kind == B2P_SCHED_GUIDANCE_DDPM || L.sc.kind == B2P_SCHED_INPAINT_DDPM);
  a.inpaint = (L.sc.kind == B2P_SCHED_INPAINT_DDIM || L.sc.kind == B2P_SCHED_INPAINT_DDPM);
  a.pred = L.sc.prediction_type;
  a.clip_mode = L.sc.thresholding ? (L.sc.sample_max_value == 1.0f ? 2 : 3) : (L.sc.clip_sample ? 1 : 0);
  a.clip_range = L.sc.clip_sample_range;
  a.noise_on = L.k.t > 0;
  a.use_clipped = (L.flags & B2P_STEP_USE_CLIPPED_OUTPUT) ? 1 : 0;
  a.eta = L.eta; a.magic = L.magic; a.flags = L.flags; a.k = L.k;
  bool needs_noise = (a.ddpm && a.noise_on) || (a.inpaint && a.traj && a.mask && a.noise_on) || (!a.ddpm && a.eta > 0.f);
  if (needs_noise && !L.noise && !L.seed) return B2P_ERR_INVALID_ARG;
  a.seed = (needs_noise && !L.noise) ? L.seed : nullptr;
  a.noise_step = L.noise_step;
  *out = a;
  return B2P_OK;
}

int launch_sched_step(const SchedLaunch& L, cudaStream_t s) {
  if (!L.mo || !L.prev) return B2P_ERR_INVALID_ARG;
  SchedK a;
  if (int rc = sched_make_args(L, &a)) return rc;
  const int n = a.n;
  float* thr = nullptr;
  if (a.clip_mode == 3) {
    // inside a plan the handle lends a scratch vector (no allocation node in the captured graph); the stand-alone step entry
    // has no handle and takes a stream-ordered allocation
    float* ts = L.thr_scratch;
    if (!ts) { B2P_CUDA_TRY(cudaMallocAsync((void**)&thr, sizeof(float) * L.B, s)); ts = thr; }
    threshold_s_kernel<<<L.B, 128, sizeof(float) * a.HD, s>>>(a, L.sc.dynamic_thresholding_ratio, L.sc.sample_max_value, ts);
    a.thr_s = ts;
  }
  int n4 = n / 4;
  cudaError_t le = launch_pdl(sched_step_kernel, dim3((n4 + 255) / 256), dim3(256), 0, s, a);
  if (le != cudaSuccess) return (int)le;
  if (thr) B2P_CUDA_TRY(cudaFreeAsync(thr, s));
  return (int)cudaGetLastError();
}

}  // namespace b2p

// ------------------------------------------------------------------------------------------------------------
// Host-side scalar arithmetic of the diffusers base classes (restated 0.28.0), fp32, in the
// reference's operation order.  Built with -ffp-contract=off.
// ------------------------------------------------------------------------------------------------------------
extern "C" int b2p_alphas_cumprod(const char* schedule, int32_t n, float beta_start, float beta_end, float* out) {
  if (!schedule || !out || n <= 0) return B2P_ERR_INVALID_ARG;
  std::vector<float> betas(n);
  std::string s(schedule);
  if (s == "squaredcos_cap_v2") {
    auto bar = [](double u) { double c = cos((u + 0.008) / 1.008 * M_PI / 2); return c * c; };
    for (int i = 0; i < n; ++i) {
      double b = 1.0 - bar((double)(i + 1) / n) / bar((double)i / n);
      betas[i] = (float)(b < 0.999 ? b : 0.999);
    }
  } else if (s == "linear" || s == "scaled_linear") {
    // torch.linspace(start, end, n, dtype=float32): step = (end-start)/(n-1) in fp32, symmetric fill from both ends.  torch's CPU
    // kernel is vectorised (base + step * lane), so its last bit depends on the host ISA: this scalar form agrees to 1 ulp
    // (tests/test_abi.py).  The shipped schedule (squaredcos_cap_v2) is exact; the Python scheduler classes use torch's betas.
    float a = beta_start, b = beta_end;
    if (s == "scaled_linear") { a = (float)sqrt((double)beta_start); b = (float)sqrt((double)beta_end); }
    float step = n > 1 ? (b - a) / (float)(n - 1) : 0.f;
    int half = n / 2;
    for (int i = 0; i < n; ++i) betas[i] = (i < half) ? a + step * (float)i : b - step * (float)(n - 1 - i);
    if (s == "scaled_linear") for (int i = 0; i < n; ++i) betas[i] = betas[i] * betas[i];
  } else {
    return B2P_ERR_INVALID_ARG;
  }
  // torch.cumprod on CPU accumulates fp32 inputs in double (at::acc_type) and rounds every output to fp32
  double p = 1.0;
  for (int i = 0; i < n; ++i) { float al = 1.0f - betas[i]; p *= (double)al; out[i] = (float)p; }
  return B2P_OK;
}

extern "C" int b2p_timesteps(int32_t n_train, int32_t n_inf, int64_t* out) {
  if (!out || n_inf <= 0 || n_inf > n_train) return B2P_ERR_INVALID_ARG;
  int64_t ratio = n_train / n_inf;
  for (int i = 0; i < n_inf; ++i) out[i] = (int64_t)(n_inf - 1 - i) * ratio;
  return B2P_OK;
}

extern "C" int b2p_step_coeffs_compute(const b2p_sched_config* sc, const float* ac, int32_t n_inf, int32_t t, float eta,
                                       b2p_step_coeffs* o) {
  if (!sc || !ac || !o || n_inf <= 0 || t < 0 || t >= sc->num_train_timesteps) return B2P_ERR_INVALID_ARG;
  memset(o, 0, sizeof(*o));
  int p = t - sc->num_train_timesteps / n_inf;
  float a_t = ac[t];
  float a_p = p >= 0 ? ac[p] : 1.0f;
  float b_t = 1.0f - a_t, b_p = 1.0f - a_p;
  o->t = t; o->t_prev = p; o->alpha_prod_t = a_t; o->alpha_prod_t_prev = a_p;
  o->sqrt_alpha_prod_t = sqrtf(a_t); o->sqrt_beta_prod_t = sqrtf(b_t);
  o->sqrt_alpha_prod_t_prev = sqrtf(a_p);
  o->sqrt_one_minus_alpha_prod_t_prev = sqrtf(1.0f - a_p);
  float cur_a = a_t / a_p;
  float variance = (b_p / b_t) * (1.0f - cur_a);
  bool ddpm = (sc->kind == B2P_SCHED_GUIDANCE_DDPM || sc->kind == B2P_SCHED_INPAINT_DDPM);
  if (ddpm) {
    if (variance < 1e-20f) variance = 1e-20f;
    o->variance = variance;
    o->std_dev_t = sqrtf(variance);
    float cur_b = 1.0f - cur_a;
    o->x0_coeff = (sqrtf(a_p) * cur_b) / b_t;
    o->sample_coeff = sqrtf(cur_a) * b_p / b_t;
  } else {
    o->variance = variance;
    o->std_dev_t = eta * sqrtf(variance);
    o->dir_coeff = sqrtf(1.0f - a_p - o->std_dev_t * o->std_dev_t);
  }
  o->guidance_grad_scale = expf(0.5f * o->variance);
  return B2P_OK;
}

extern "C" int b2p_philox_normal(uint64_t key, int32_t steps, int64_t n_per_step, float* out, void* stream) {
  if (!out || steps <= 0 || n_per_step <= 0 || n_per_step % 4 != 0 || n_per_step / 4 > 0x7fffffff) return B2P_ERR_INVALID_ARG;
  const int groups = (int)(n_per_step / 4);
  b2p::philox_fill_kernel<<<dim3((groups + 255) / 256, steps), 256, 0, (cudaStream_t)stream>>>(key, groups, steps, out);
  return (int)cudaGetLastError();
}

extern "C" int b2p_sched_step(const b2p_sched_config* sc, const b2p_step_coeffs* k, const float* model_output,
                              const float* model_output_uncond, float cfg_scale, const float* sample, const float* noise,
                              const float* target_traj, const float* target_mask, float* prev_out, float* x0_out,
                              int32_t B, int32_t H, int32_t D, float eta, float magic_num, int32_t flags, void* stream) {
  if (!sc || !k) return B2P_ERR_INVALID_ARG;
  b2p::SchedLaunch L{*sc, *k, model_output, model_output_uncond, cfg_scale, sample, noise, target_traj, target_mask,
                     prev_out, x0_out, B, H, D, eta, magic_num, flags, nullptr, 0u, nullptr};
  return b2p::launch_sched_step(L, (cudaStream_t)stream);
}
