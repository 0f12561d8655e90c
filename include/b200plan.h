/*
 * b200plan.h — C ABI of the B200-native diffusion-planning hot path.
 *
 * The reference (Justin900429/autonomous_driving_with_diffusion_model) is pure Python/PyTorch and has no FFI of its
 * own; this ABI is what the Python drop-in layer (autonomous_driving_with_diffusion_model_b200/*.py) binds with
 * ctypes.  Each entry point cites the reference interface it replaces.  Conventions:
 *   - plain pointers and sizes only; every tensor is fp32, contiguous, in the reference's own layout
 *     (trajectories [B, H, D]; features [B, dim]; timesteps int64);
 *   - unless a function name ends in _host, data pointers are DEVICE pointers and work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) with no hidden synchronisation, so calls are
 *     CUDA-graph capturable;
 *   - return value: 0 = ok, <0 = b2p_status below, >0 = a cudaError_t; nothing throws across the boundary;
 *   - a handle is bound to one device, owns packed weights / tables / workspace / CUDA graphs and is not thread-safe;
 *     distinct handles are independent.  There is no CPU fallback.
 */
#ifndef B200PLAN_H_
#define B200PLAN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define B2P_ABI_VERSION 8

typedef enum {
  B2P_OK = 0,
  B2P_ERR_INVALID_ARG = -1,   /* bad shape / null pointer / unsupported configuration            */
  B2P_ERR_UNKNOWN_WEIGHT = -2,/* state_dict key is not part of the configured model              */
  B2P_ERR_BAD_SHAPE = -3,     /* numel does not match the key's shape                            */
  B2P_ERR_NOT_FINALIZED = -4, /* forward/plan called before b2p_finalize_weights                 */
  B2P_ERR_MISSING_WEIGHT = -5,/* b2p_finalize_weights called with keys still unset               */
  B2P_ERR_NO_DEVICE = -6,     /* no usable CUDA device / not an sm_100 part                      */
  B2P_ERR_STATE = -7          /* call sequence error (e.g. plan before set_timesteps)            */
} b2p_status;

/* GuidanceType, misc/constant.py:17-20 */
typedef enum { B2P_NO_GUIDANCE = 0, B2P_FREE_GUIDANCE = 1, B2P_CLASSIFIER_GUIDANCE = 2 } b2p_guidance_type;

/* scheduler family, scheduler/__init__.py:6-11 */
typedef enum {
  B2P_SCHED_GUIDANCE_DDIM = 0,   /* scheduler/guidance_ddim_scheduler.py   */
  B2P_SCHED_GUIDANCE_DDPM = 1,   /* scheduler/guidance_ddpm_scheduler.py   */
  B2P_SCHED_INPAINT_DDIM = 2,    /* scheduler/inpainting_ddim_scheduler.py */
  B2P_SCHED_INPAINT_DDPM = 3     /* scheduler/inpainting_ddpm_scheduler.py */
} b2p_sched_kind;

typedef enum { B2P_PRED_EPSILON = 0, B2P_PRED_SAMPLE = 1, B2P_PRED_V = 2 } b2p_pred_type;
typedef enum { B2P_BETA_SQUAREDCOS_CAP_V2 = 0, B2P_BETA_LINEAR = 1, B2P_BETA_SCALED_LINEAR = 2 } b2p_beta_schedule;

/* arithmetic of the conv contractions */
typedef enum {
  B2P_PREC_FP32 = 0,     /* CUDA-core FFMA, fp32 everywhere (parity mode, <=1e-3 vs the reference)              */
  B2P_PREC_BF16X3 = 1,   /* tcgen05 bf16 hi/lo split (3 MMAs), fp32 accumulate: fp32-class parity on tensor cores */
  B2P_PREC_BF16 = 2      /* tcgen05 single-pass bf16 operands, fp32 accumulate/GroupNorm/state (throughput mode)  */
} b2p_precision;

/* Constructor arguments of TemporalMapUnet (modeling/temporal.py:59-68) + build_model (modeling/temporal.py:248-258). */
typedef struct {
  int32_t horizon;         /* cfg.MODEL.HORIZON, 16                         */
  int32_t transition_dim;  /* cfg.MODEL.TRANSITION_DIM, 7                   */
  int32_t dim;             /* cfg.MODEL.DIM, 64                             */
  int32_t n_mults;         /* len(cfg.MODEL.DIM_MULTS), 4                   */
  int32_t dim_mults[8];    /* (1, 2, 4, 8)                                  */
  int32_t guidance;        /* b2p_guidance_type (TRAIN.USE_COND)            */
  int32_t precision;       /* b2p_precision                                 */
} b2p_model_config;

/* diffusers-style scheduler config as constructed at interact.py:81-94 */
typedef struct {
  int32_t kind;                 /* b2p_sched_kind                                                   */
  int32_t num_train_timesteps;  /* TRAIN.SAMPLE_STEPS, 100                                          */
  int32_t prediction_type;      /* b2p_pred_type; "sample" in every shipped config                  */
  int32_t thresholding;         /* interact.py:88 passes True                                       */
  int32_t clip_sample;          /* diffusers default True (only used when thresholding == 0)        */
  float clip_sample_range;      /* 1.0                                                              */
  float dynamic_thresholding_ratio; /* 0.995                                                        */
  float sample_max_value;       /* 1.0  (=> thresholding == clamp(-1, 1))                           */
  int32_t beta_schedule;        /* b2p_beta_schedule; TRAIN.NOISE_SCHEDULER.TYPE = squaredcos_cap_v2 */
  float beta_start, beta_end;   /* only read by the linear schedules (1e-4, 0.02)                   */
} b2p_sched_config;

/* One scheduler step's scalar coefficients, all derived in fp32 in the reference's operation order. */
typedef struct {
  int32_t t, t_prev;
  float alpha_prod_t, alpha_prod_t_prev;
  float sqrt_alpha_prod_t, sqrt_beta_prod_t;        /* a_t**0.5, (1-a_t)**0.5                       */
  float sqrt_alpha_prod_t_prev;                     /* a_prev**0.5                                  */
  float sqrt_one_minus_alpha_prod_t_prev;           /* (1-a_prev)**0.5 (inpainting known part)      */
  float variance;                                   /* _get_variance(t[, prev])                     */
  float std_dev_t;                                  /* DDIM: eta*variance**0.5; DDPM: variance**0.5 */
  float dir_coeff;                                  /* DDIM: (1 - a_prev - std_dev_t**2)**0.5       */
  float x0_coeff, sample_coeff;                     /* DDPM formula (7) coefficients                */
  float guidance_grad_scale;                        /* exp(0.5*variance) (guidance_ddim_scheduler.py:91) */
} b2p_step_coeffs;

typedef struct b2p_handle_s* b2p_handle;

/* ---- library ---------------------------------------------------------------------------------------------- */
int b2p_abi_version(void);
const char* b2p_status_string(int status);
/* last error message recorded on this handle ("" if none) */
const char* b2p_last_error(b2p_handle h);

/* ---- model lifetime: replaces build_model(cfg).to(device) (modeling/temporal.py:248, interact.py:101) ------- */
int b2p_create(const b2p_model_config* cfg, int device, b2p_handle* out);
int b2p_destroy(b2p_handle h);
/* replaces model.load_state_dict(...) (interact.py:104): one call per UNet state_dict key (reference key names,
 * SURVEY.md Appendix A; `perception.*` keys are not part of this ABI).  `data` is a HOST fp32 pointer. */
int b2p_load_weight(b2p_handle h, const char* key, const float* data, int64_t numel);
/* number of keys the configured model expects; key i and its numel (for iteration from the host language) */
int b2p_num_weights(b2p_handle h);
int b2p_weight_info(b2p_handle h, int index, const char** key, int64_t* numel);
/* pack + upload; must be called after the last b2p_load_weight and again after any weight change */
int b2p_finalize_weights(b2p_handle h);
int b2p_set_precision(b2p_handle h, int precision);
/* Denoiser evaluations of at most `max_samples` trajectories (classifier-free doubling included) run the small-batch
 * exact-fp32 GEMV kernels whatever the precision mode (closed-loop planning calls generate_traj with one trajectory,
 * carla_agent; default B2P_SMALL_BATCH_DEFAULT).  0 disables that path. */
#define B2P_SMALL_BATCH_DEFAULT 4
int b2p_set_small_batch_max(b2p_handle h, int max_samples);
/* Developer switch (default on; environment B2P_CHAIN=0 turns it off for new handles): in the tensor-core precisions the
 * 64-channel layers at the full-resolution end of the U-Net run as row-owned CHAINS (one launch for downs.0, one for the
 * tail of ups.<last> + final_conv, fused with the scheduler step between two evaluations of a plan without guidance).
 * 0 = every layer is its own launch.  Same results up to fp32 summation order.                                     */
int b2p_set_chain(b2p_handle h, int enabled);

/* ---- camera frame -> encoder input: replaces T.ToTensor() + T.Normalize(mean, std) (interact.py:72-77, 170-172) for
 * uint8 frames already on the device.  frames [N,H,W,3] uint8 (4-byte aligned), out [N,H,W,3] fp32 (16-byte aligned; the
 * channels-last memory of the logical [N,3,H,W] tensor), n_pixels = N*H*W.  out = ((u8 / 255) - mean[c]) / std[c],
 * bit-identical to the host transform.  No handle: the transform has no state. */
int b2p_preprocess_frames(const uint8_t* frames_nhwc, float* out_nhwc, int64_t n_pixels, const float mean[3], const float std_[3],
                          void* stream);

/* ---- image-encoder stem (bf16 mode of the encoder): replaces conv1 + bn1 + relu and maxpool of the reference ResNet
 * (modeling/resnet.py:193-198, 279-282) on the device.  img: fp32, logical [N,3,H,W] with ELEMENT strides
 * (stride_n, stride_c, stride_h, stride_w) — NCHW and channels-last both work.  weight_image: the BatchNorm-folded conv1 weights as
 * a 24,576-byte bf16 operand image: K index k = (kernel_row * 7 + kernel_col) * 3 + channel, padded to 192; element (n, k) at byte
 * (k / 64) * 8192 + n * 128 + ((((k % 64) / 8) ^ (n % 8)) * 16) + (k % 8) * 2  (K-major, 128-byte swizzle).  bias [64] fp32 (folded).
 * out: bf16 [N, OH, OW, 64] (channels-last), OH = (H - 1) / 2 + 1, OW likewise, 16-byte aligned.  bf16 products, fp32 accumulation. */
int b2p_encoder_stem_bf16(const float* img, int64_t stride_n, int64_t stride_c, int64_t stride_h, int64_t stride_w, int32_t N, int32_t H,
                          int32_t W, const void* weight_image, const float* bias, void* out_nhwc_bf16, void* stream);
/* conv1 + bn1 + relu + maxpool(3, 2, 1) in one kernel (modeling/resnet.py:279-282): same arguments as b2p_encoder_stem_bf16, but out is the POOLED
 * tensor bf16 [N, PH, PW, 64] with PH = ((H - 1) / 2) / 2 + 1, PW likewise; bit-identical to b2p_encoder_stem_bf16 followed by
 * b2p_maxpool3x3s2_nhwc_bf16 (conv1's 1.9 GB output for 256 frames never reaches HBM).  Strides must be non-negative. */
int b2p_encoder_stem_pool_bf16(const float* img, int64_t stride_n, int64_t stride_c, int64_t stride_h, int64_t stride_w, int32_t N, int32_t H,
                               int32_t W, const void* weight_image, const float* bias, void* out_nhwc_bf16, void* stream);
/* MaxPool2d(kernel 3, stride 2, padding 1) on bf16 [N,H,W,C] -> [N,(H-1)/2+1,(W-1)/2+1,C], C a multiple of 8, 16-byte aligned. */
int b2p_maxpool3x3s2_nhwc_bf16(const void* in, void* out, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);
/* ---- image-encoder body (bf16 mode): one convolution of ResNet-34's layer1..layer4 with its BatchNorm folded in, the residual add and the
 * ReLU (modeling/resnet.py:56-102: conv3x3 -> bn -> relu -> conv3x3 -> bn -> (+ identity | downsample(x)) -> relu; 199-214, 283-286).
 * in: bf16 [N,H,W,Cin] (channels-last);  ksize/stride: 3/1 (pad 1), 3/2 (pad 1) or 1/2 (pad 0, the projection shortcut);
 * w_packed: bf16 [ksize*ksize][Cout][Cin] = weight.permute(2,3,0,1) with the BatchNorm scale folded in;  bias: fp32 [Cout] (folded);
 * res: bf16 [N,OH,OW,Cout] added before the ReLU, or NULL;  out: bf16 [N,OH,OW,Cout], OH = (H-1)/stride+1, OW likewise;  relu: 0/1.
 * Cin, Cout multiples of 64; all pointers 16-byte aligned.  bf16 products, fp32 accumulation, one rounding
 * of the result to bf16. */
int b2p_encoder_conv_bf16(const void* in_nhwc, int32_t N, int32_t H, int32_t W, int32_t Cin, const void* w_packed, const float* bias, const void* res_nhwc,
                          void* out_nhwc, int32_t Cout, int32_t ksize, int32_t stride, int32_t relu, void* stream);

/* ---- denoiser: replaces TemporalMapUnet.forward (modeling/temporal.py:197-245) with the image feature hoisted --
 * x        [B, H, D]            noisy trajectories
 * feat     [feat_rows, dim]     perception(img) (modeling/temporal.py:203); feat_rows in {B, B/2 (CFG repeat)}
 * t        [t_count] int64      t_count in {1, B, B/2}; broadcast / repeated as temporal.py:206-211 does
 * cond     [B, 2] or NULL       FREE_GUIDANCE target-point condition (NULL == zeros, temporal.py:207)
 * out      [B, H, D]            NO/FREE: model output.  CLASSIFIER: cat[state, action] (temporal.py:237-241)
 * action_out [B, H, 3], time_embed_out [B, dim]: CLASSIFIER only (return_action_and_time_only, temporal.py:235-236);
 *          may be NULL.  If out == NULL in CLASSIFIER mode the state predictor is skipped.                        */
int b2p_unet_forward(b2p_handle h, const float* x, const float* feat, int32_t feat_rows, const int64_t* t,
                     int32_t t_count, const float* cond, float* out, float* action_out, float* time_embed_out,
                     int32_t B, void* stream);

/* ---- TrajPredict: replaces model.state_pred(action[:, :-1], time_embed) (modeling/helpers.py:53-59, interact.py:158)
 * action [B, H, 3] (rows 0..H-2 are used), time_embed [B, dim] -> state [B, H-1, D-3]                             */
int b2p_state_pred(b2p_handle h, const float* action, const float* time_embed, float* state, int32_t B, void* stream);

/* vector-Jacobian product of the call above wrt `action` (what torch.autograd.grad(loss, [x_guidance, action]) needs,
 * control/guidance.py:45-48): grad_state [B, H-1, D-3] -> grad_action [B, H, 3] (row H-1 is zero)                   */
int b2p_state_pred_vjp(b2p_handle h, const float* action, const float* time_embed, const float* grad_state,
                       float* grad_action, int32_t B, void* stream);

/* ---- classifier guidance: replaces GuidanceLoss.forward + TargetGuidance (control/guidance.py:35-59,
 * control/guidance_loss.py:10-22) for GUIDANCE.STEP == 1, batched as the per-sample map of the B=1 rule.
 * model_output [B,H,D] = cat[state, action] is updated IN PLACE (as the reference's .detach() alias does);
 * the action gradient is the analytic VJP through TrajPredict.                                                     */
int b2p_classifier_guidance(b2p_handle h, float* model_output, const float* time_embed, const float* target,
                            float grad_scale, float classifier_scale, int32_t B, void* stream);

/* ---- scheduler: replaces {Guidance,Inpainting}{DDIM,DDPM}Scheduler.step (scheduler/*.py) -------------------- */
/* host-only: betas/alphas_cumprod of the diffusers base (restated 0.28.0), beta_schedule in
 * {"squaredcos_cap_v2","linear","scaled_linear"}; writes num_train_timesteps floats */
int b2p_alphas_cumprod(const char* beta_schedule, int32_t num_train_timesteps, float beta_start, float beta_end, float* out);
/* host-only: "leading" timesteps of set_timesteps(); writes num_inference_steps int64 */
int b2p_timesteps(int32_t num_train_timesteps, int32_t num_inference_steps, int64_t* out);
/* host-only: coefficients for one step */
int b2p_step_coeffs_compute(const b2p_sched_config* sc, const float* alphas_cumprod, int32_t num_inference_steps,
                            int32_t t, float eta, b2p_step_coeffs* out);

/* flags for b2p_sched_step */
#define B2P_STEP_ZERO_FIRST_WAYPOINT 1  /* trajs[:, 0, :3] = 0 after the step (interact.py:164)                    */
#define B2P_STEP_FINAL_POSTPROCESS 2    /* clamp(-1,1) and [..., :2] *= magic_num (interact.py:166-167)            */
#define B2P_STEP_USE_CLIPPED_OUTPUT 4   /* use_clipped_model_output=True (DDIM variants)                           */

/* One fused elementwise launch.  All tensors [B, H, D] device fp32.
 * model_output_uncond != NULL: classifier-free mix  u + cfg_scale*(c - u) (interact.py:142-144) with model_output = c.
 * noise may be NULL when the step does not consume noise; target_traj/target_mask NULL => no inpainting blend.
 * x0_out (pred_original_sample) may be NULL.  prev_out may alias sample.                                          */
int b2p_sched_step(const b2p_sched_config* sc, const b2p_step_coeffs* k, const float* model_output,
                   const float* model_output_uncond, float cfg_scale, const float* sample, const float* noise,
                   const float* target_traj, const float* target_mask, float* prev_out, float* x0_out,
                   int32_t B, int32_t H, int32_t D, float eta, float magic_num, int32_t flags, void* stream);

/* ---- whole plan: replaces Agent.generate_traj (interact.py:115-168 == e2e_driving/diffusion_agent.py:179-232) --
 * The T-step loop (denoiser + CFG mix / classifier guidance + scheduler step + waypoint overwrite + post-process) is
 * captured once per (B, T, scheduler) as a CUDA graph and replayed.                                               */
typedef struct {
  b2p_sched_config sched;
  int32_t num_inference_steps;  /* EVAL.SAMPLE_STEPS                                  */
  float eta;                    /* callers never pass eta => 0                        */
  float free_scale;             /* GUIDANCE.FREE_SCALE                                */
  float classifier_scale;       /* GUIDANCE.CLASSIFIER_SCALE                          */
  float magic_num;              /* model.magic_num, 23.315 (modeling/temporal.py:195) */
  int32_t postprocess;          /* 1: clamp + scale as interact.py:166-167            */
  int32_t use_graph;            /* 1: capture/replay a CUDA graph                     */
} b2p_plan_config;

/* device-pointer variant.  x_init [B,H,D]; feat [B,dim]; target [B,2] or NULL; noise [T,B,H,D] or NULL;
 * target_traj/target_mask [B,H,D] or NULL; out [B,H,D].  A scheduler that consumes noise (DDPM, inpainting blend, eta > 0)
 * and gets noise == NULL draws it inside the scheduler kernel (Philox4x32-10; replaces randn_tensor at
 * guidance_ddpm_scheduler.py:154-157), keyed by b2p_set_noise_seed and a per-plan counter: statistically equivalent, not
 * the same stream as torch's generator — parity runs inject `noise`.                                              */
int b2p_plan(b2p_handle h, const b2p_plan_config* pc, const float* x_init, const float* feat, const float* target,
             const float* noise, const float* target_traj, const float* target_mask, float* out, int32_t B,
             void* stream);
/* host-pointer variant (pinned or pageable host memory): H2D of the inputs, the loop, D2H of the result, and a
 * stream synchronise before returning.  This is the end-to-end entry a non-PyTorch caller uses.                   */
int b2p_plan_host(b2p_handle h, const b2p_plan_config* pc, const float* x_init, const float* feat, const float* target,
                  const float* noise, const float* target_traj, const float* target_mask, float* out, int32_t B);

/* the same without the final synchronise: H2D copies, the loop and the D2H copy are enqueued on the handle's private stream
 * and the call returns; b2p_sync(h) joins.  `noise_batch` is the batch extent of the caller's noise tensor
 * [T, noise_batch, H, D] when this call plans the slice starting at `noise` (0 or B = dense).  Host buffers should be
 * pinned (pageable memory makes the copies synchronous) and must stay valid until b2p_sync returns.                */
int b2p_plan_host_async(b2p_handle h, const b2p_plan_config* pc, const float* x_init, const float* feat, const float* target,
                        const float* noise, int32_t noise_batch, const float* target_traj, const float* target_mask, float* out,
                        int32_t B);
int b2p_sync(b2p_handle h);
/* multi-GPU entry (north_star "independent planning requests are sharded by batch across the 8 GPUs of one
 * box with no NCCL on the sampling path"): one handle per device (same weights loaded into each), contiguous batch split
 * (the first B % n shards take one extra trajectory), all shards enqueued, then joined.  Host tensors as b2p_plan_host.
 * Replaces the single caller / single batch of interact.py:115-168 for a fleet-sized batch.                          */
int b2p_plan_sharded_host(const b2p_handle* handles, int32_t n_handles, const b2p_plan_config* pc, const float* x_init,
                          const float* feat, const float* target, const float* noise, const float* target_traj,
                          const float* target_mask, float* out, int32_t B);

/* seed of the in-kernel noise (resets the per-plan counter): the same seed and call sequence reproduce the same plans */
int b2p_set_noise_seed(b2p_handle h, uint64_t seed);
/* Philox key the last plan with in-kernel noise used, and the noise tensor that key stands for: out [steps, n_per_step]
 * (n_per_step = B*H*D, device fp32, 16-byte aligned) — injecting it as `noise` reproduces the plan bit for bit.     */
uint64_t b2p_last_noise_key(b2p_handle h);
int b2p_philox_normal(uint64_t key, int32_t steps, int64_t n_per_step, float* out, void* stream);

/* ---- plan post-processing -> control for a fleet of vehicles ------------------------------- */
/* PID.* and CONTROL.* of the reference configuration (config.py:67-86) */
typedef struct {
  double turn_kp, turn_ki, turn_kd; int32_t turn_n;      /* PID.TURN_*  (window length <= 127)  */
  double speed_kp, speed_ki, speed_kd; int32_t speed_n;  /* PID.SPEED_*                          */
  double aim_dist, angle_thresh, dist_thresh, brake_speed, brake_ratio, clip_delta, max_throttle;  /* CONTROL.* */
} b2p_control_config;
/* bytes of device memory holding the PID windows of V vehicles (the per-vehicle deques of control/pid.py:10) */
int64_t b2p_fleet_state_bytes(const b2p_control_config* c, int32_t V);
/* zero the windows (a fresh Controller per vehicle, control/controller.py:8-21) */
int b2p_fleet_reset(const b2p_control_config* c, void* state, int32_t V, void* stream);
/* one control tick for every vehicle: replaces Controller.control_pid (control/controller.py:29-76) + PIDController.step
 * (control/pid.py:16-28) called in a Python loop.  waypoints [V,N,2] (ego frame, metres), velocity [V], target [V,2] ->
 * out [V,3] = (throttle, steer, brake in {0,1}); float64 arithmetic, the windows advance by one.                     */
int b2p_fleet_control_pid(const b2p_control_config* c, void* state, const float* waypoints, int32_t N, const float* velocity,
                          const float* target, float* out, int32_t V, void* stream);
/* replaces post_process_control on traj[:, 0, -3:] (interact.py:218-229, 296-297): trajs [B,H,D] -> out [B,3] */
int b2p_fleet_post_process(const float* trajs, float* out, int32_t B, int32_t H, int32_t D, void* stream);

/* ---- introspection for tests / bench ---------------------------------------------------------------------- */
/* number of kernel launches the last b2p_unet_forward / b2p_plan call enqueued (graph replay counts its nodes) */
int64_t b2p_last_launch_count(b2p_handle h);
/* nominal FLOPs (2*MAC, padding taps counted, as torch FlopCounterMode) of one denoiser evaluation per trajectory */
int64_t b2p_unet_flops_per_sample(b2p_handle h);
/* bytes of packed weights resident on the device */
int64_t b2p_weight_bytes(b2p_handle h);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* B200PLAN_H_ */
