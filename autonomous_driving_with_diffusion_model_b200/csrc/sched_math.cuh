// Arithmetic of one fused scheduler step, shared by sched.cu (stand-alone launch) and chain64.cu (fused behind the denoiser's
// head inside the seam kernel).  Round-to-nearest intrinsics in the reference's operation order (no FMA contraction), so
// both users are bit-identical to the fp32 CPU oracle given the same model output.
// Replaces scheduler/guidance_ddim_scheduler.py:60-173, guidance_ddpm_scheduler.py:59-178, inpainting_ddim_scheduler.py:10-153,
// inpainting_ddpm_scheduler.py:10-146, interact.py:142-144,164,166-167.
#pragma once
#include "common.cuh"

namespace b2p {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

struct SchedK {
  const float* mo; const float* mo_u; float cfg_scale;
  const float* sample; const float* noise; const float* traj; const float* mask;
  float* prev; float* x0_out;
  const float* thr_s;  // per-sample dynamic threshold s (clip_mode 3)
  int n, HD, D;
  int ddpm, inpaint, pred, clip_mode;  // clip_mode: 0 none, 1 clamp(+-range), 2 threshold with s==1, 3 dynamic
  float clip_range;
  int noise_on;       // t > 0
  int use_clipped;
  float eta, magic; int flags;
  const unsigned long long* seed; unsigned noise_step;
  b2p_step_coeffs k;
};

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so a thread's draw depends only on (seed, element group, step)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
// four independent standard normals (Box-Muller on 24-bit uniforms; u1 in (0,1] so the log is finite)
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned group, unsigned step) {
  const uint4 r = philox4x32_10(make_uint4(group, step, 0x6e6f6973u, 0x65u), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float s24 = 1.0f / 16777216.0f;
  const float u0 = (float)((r.x >> 8) + 1u) * s24, u1 = (float)(r.y >> 8) * s24;
  const float u2 = (float)((r.z >> 8) + 1u) * s24, u3 = (float)(r.w >> 8) * s24;
  const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float sa, ca, sb, cb;
  sincospif(2.0f * u1, &sa, &ca);
  sincospif(2.0f * u3, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

__device__ __forceinline__ float x0_of(const SchedK& a, float m, float x) {
  if (a.pred == B2P_PRED_SAMPLE) return m;
  if (a.pred == B2P_PRED_EPSILON) return dvd(sub(x, mul(a.k.sqrt_beta_prod_t, m)), a.k.sqrt_alpha_prod_t);
  return sub(mul(a.k.sqrt_alpha_prod_t, x), mul(a.k.sqrt_beta_prod_t, m));
}

// sample = index of the trajectory, pos = offset inside its [H*D] block, col = pos % D (tracked by the caller: one division per
// thread instead of three per element)
__device__ __forceinline__ float step_one(const SchedK& a, int sample, int pos, int col, float m, float mu, float x, float nz, float tj, float mk,
                                          float* x0_store) {
  if (a.mo_u) m = add(mu, mul(a.cfg_scale, sub(m, mu)));  // u + s*(c - u)
  float x0 = x0_of(a, m, x);
  float eps = 0.f;
  if (!a.ddpm) {
    if (a.pred == B2P_PRED_SAMPLE) eps = dvd(sub(x, mul(a.k.sqrt_alpha_prod_t, x0)), a.k.sqrt_beta_prod_t);  // un-clamped x0
    else if (a.pred == B2P_PRED_EPSILON) eps = m;
    else eps = add(mul(a.k.sqrt_alpha_prod_t, m), mul(a.k.sqrt_beta_prod_t, x));
  }
  if (a.clip_mode == 1) x0 = clampf(x0, -a.clip_range, a.clip_range);
  else if (a.clip_mode == 2) x0 = dvd(clampf(x0, -1.f, 1.f), 1.f);
  else if (a.clip_mode == 3) { float s = a.thr_s[sample]; x0 = dvd(clampf(x0, -s, s), s); }
  *x0_store = x0;
  float prev;
  if (!a.ddpm) {
    if (a.use_clipped) eps = dvd(sub(x, mul(a.k.sqrt_alpha_prod_t, x0)), a.k.sqrt_beta_prod_t);
    float dir = mul(a.k.dir_coeff, eps);
    prev = add(mul(a.k.sqrt_alpha_prod_t_prev, x0), dir);
    if (a.inpaint) prev = add(prev, a.k.variance);  // quirk: scalar sigma_t^2 added as an offset
  } else {
    prev = add(mul(a.k.x0_coeff, x0), mul(a.k.sample_coeff, x));
    float var_term = a.noise_on ? mul(a.k.std_dev_t, nz) : 0.f;
    prev = add(prev, var_term);
  }
  if (a.inpaint && a.traj && a.mask) {
    float kn = a.noise_on ? nz : 0.f;
    float known = add(mul(a.k.sqrt_alpha_prod_t_prev, tj), mul(a.k.sqrt_one_minus_alpha_prod_t_prev, kn));
    prev = add(mul(mk, known), mul(sub(1.0f, mk), prev));
  }
  if (!a.ddpm && a.eta > 0.f) prev = add(prev, mul(a.k.std_dev_t, nz));
  if ((a.flags & B2P_STEP_ZERO_FIRST_WAYPOINT) && pos < 3) prev = 0.f;
  if (a.flags & B2P_STEP_FINAL_POSTPROCESS) {
    prev = clampf(prev, -1.f, 1.f);
    if (col < 2) prev = mul(prev, a.magic);
  }
  return prev;
}

int sched_make_args(const SchedLaunch& L, SchedK* out);   // sched.cu

}  // namespace b2p
