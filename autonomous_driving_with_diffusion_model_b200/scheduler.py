"""Drop-in diffusers-style schedulers whose ``step()`` is ONE fused CUDA launch through the C ABI.

Mirrors scheduler/__init__.py:6-11 of the reference: ``GuidanceDDIMScheduler``, ``GuidanceDDPMScheduler``,
``InpaintingDDIMScheduler``, ``InpaintingDDPMScheduler`` with the constructor keywords interact.py:81-94 passes, the
diffusers surface the callers touch (``set_timesteps``, ``timesteps``, ``step`` -> ``.prev_sample`` /
``.pred_original_sample``, ``config``, ``alphas_cumprod``, ``num_inference_steps``, ``add_noise``) and the reference's
error messages.  All coefficient arithmetic happens in libb200plan (fp32, reference operation order); nothing here
depends on the third-party ``diffusers`` package.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .constant import GuidanceType
from .guidance import GuidanceLoss


@dataclass
class SchedulerOutput:
    """DDIMSchedulerOutput / DDPMSchedulerOutput."""
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return (self.prev_sample, self.pred_original_sample)[i]


DDIMSchedulerOutput = DDPMSchedulerOutput = SchedulerOutput


class _FusedScheduler:
    _kind = "guidance_ddim"
    _is_ddpm = False

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True, set_alpha_to_one: bool = True,
                 steps_offset: int = 0, prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0, sample_max_value: float = 1.0,
                 timestep_spacing: str = "leading", rescale_betas_zero_snr: bool = False, variance_type: str = "fixed_small"):
        if trained_betas is not None or rescale_betas_zero_snr or timestep_spacing != "leading" or steps_offset != 0 or not set_alpha_to_one:
            raise NotImplementedError("only the configuration the reference constructs (interact.py:81-94) is supported: "
                                      "leading spacing, steps_offset 0, set_alpha_to_one, no trained_betas / zero-SNR rescale")
        if self._is_ddpm and variance_type != "fixed_small":
            raise NotImplementedError("variance_type must be 'fixed_small' (the diffusers default the reference relies on)")
        if beta_schedule not in _lib.BETA_SCHEDULES:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if prediction_type not in _lib.PRED_TYPES:
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, `sample`, or `v_prediction`")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, prediction_type=prediction_type,
                                      thresholding=thresholding, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
                                      clip_sample_range=clip_sample_range, sample_max_value=sample_max_value,
                                      timestep_spacing=timestep_spacing, rescale_betas_zero_snr=rescale_betas_zero_snr,
                                      variance_type=variance_type)
        # betas / alphas_cumprod exactly as the diffusers base builds them (torch CPU fp32, so bit-identical to what the
        # reference sees; libb200plan's b2p_alphas_cumprod is the same table for non-Python hosts)
        if beta_schedule == "squaredcos_cap_v2":
            bar = lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
            n = num_train_timesteps
            self.betas = torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), 0.999) for i in range(n)], dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            self.betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.variance_type = variance_type
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    # ---- diffusers surface -----------------------------------------------------------------------------
    def sched_config(self) -> "_lib.SchedConfig":
        c = self.config
        sc = _lib.SchedConfig()
        sc.kind = _lib.SCHED_KINDS[self._kind]
        sc.num_train_timesteps = c.num_train_timesteps
        sc.prediction_type = _lib.PRED_TYPES[c.prediction_type]
        sc.thresholding, sc.clip_sample = int(bool(c.thresholding)), int(bool(c.clip_sample))
        sc.clip_sample_range, sc.dynamic_thresholding_ratio, sc.sample_max_value = c.clip_sample_range, c.dynamic_thresholding_ratio, c.sample_max_value
        sc.beta_schedule, sc.beta_start, sc.beta_end = _lib.BETA_SCHEDULES[c.beta_schedule], c.beta_start, c.beta_end
        return sc

    def set_timesteps(self, num_inference_steps: int, device=None, timesteps=None):
        if timesteps is not None:
            raise NotImplementedError("custom timesteps are not used by the reference")
        n = self.config.num_train_timesteps
        if num_inference_steps > n:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`: {n} as the unet "
                             f"model trained with this scheduler can only handle maximal {n} timesteps.")
        ts = np.empty(num_inference_steps, dtype=np.int64)
        _lib.check(_lib.load().b2p_timesteps(n, num_inference_steps, ts.ctypes.data_as(_lib.c_int64_p)), None, "b2p_timesteps")
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def previous_timestep(self, timestep):
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def coeffs(self, timestep: int, eta: float = 0.0) -> "_lib.StepCoeffs":
        """Scalar coefficients of one step on 0-dim fp32 CPU tensors, in the reference's operation order
        (guidance_ddim_scheduler.py:86-136, guidance_ddpm_scheduler.py:92-134), handed to the kernel by value."""
        t = int(timestep)
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        p = t - self.config.num_train_timesteps // n
        ac = self.alphas_cumprod
        a_t = ac[t]
        a_p = ac[p] if p >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        k = _lib.StepCoeffs()
        k.t, k.t_prev = t, p
        k.alpha_prod_t, k.alpha_prod_t_prev = float(a_t), float(a_p)
        k.sqrt_alpha_prod_t, k.sqrt_beta_prod_t = float(a_t ** 0.5), float(b_t ** 0.5)
        k.sqrt_alpha_prod_t_prev = float(a_p ** 0.5)
        k.sqrt_one_minus_alpha_prod_t_prev = float((1.0 - a_p) ** 0.5)
        if self._is_ddpm:
            cur_a = a_t / a_p
            cur_b = 1 - cur_a
            variance = torch.clamp((1 - a_p) / (1 - a_t) * cur_b, min=1e-20)
            k.std_dev_t = float(variance ** 0.5)
            k.x0_coeff = float((a_p ** 0.5 * cur_b) / b_t)
            k.sample_coeff = float(cur_a ** 0.5 * (1 - a_p) / b_t)
        else:
            variance = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
            std = eta * variance ** 0.5
            k.std_dev_t = float(std)
            k.dir_coeff = float((1 - a_p - std ** 2) ** 0.5)
        k.variance = float(variance)
        k.guidance_grad_scale = float(torch.exp(0.5 * variance))
        return k

    def _get_variance(self, timestep, prev_timestep=None, **_):
        return torch.tensor(self.coeffs(int(timestep)).variance)

    def add_noise(self, original_samples, noise, timesteps):
        """Training-side helper (train.py:234); plain torch, not on the sampling path."""
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        a, b = ac[timesteps] ** 0.5, (1 - ac[timesteps]) ** 0.5
        while a.dim() < original_samples.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise

    # ---- the fused launch ------------------------------------------------------------------------------
    def _launch(self, model_output, timestep, sample, *, eta=0.0, use_clipped=False, noise=None, target_traj=None, target_mask=None,
                model_output_uncond=None, cfg_scale=1.0, flags=0, magic_num=23.315):
        if sample.device.type != "cuda":
            raise RuntimeError("scheduler.step runs on CUDA tensors only (no CPU fallback)")
        B, H, D = sample.shape
        if B == 0:
            return torch.empty_like(sample, dtype=torch.float32), torch.empty_like(sample, dtype=torch.float32)
        f32 = lambda t: None if t is None else t.detach().to(sample.device, torch.float32).expand(B, H, D).contiguous()  # noqa: E731
        mo, x = f32(model_output), f32(sample)
        mo_u, nz, tj, mk = f32(model_output_uncond), f32(noise), f32(target_traj), f32(target_mask)
        prev, x0 = torch.empty_like(x), torch.empty_like(x)
        sc, k = self.sched_config(), self.coeffs(int(timestep), eta)
        if use_clipped:
            flags |= _lib.STEP_USE_CLIPPED_OUTPUT
        rc = _lib.load().b2p_sched_step(C.byref(sc), C.byref(k), _lib.ptr(mo), _lib.ptr(mo_u), float(cfg_scale), _lib.ptr(x), _lib.ptr(nz),
                                        _lib.ptr(tj), _lib.ptr(mk), _lib.ptr(prev), _lib.ptr(x0), B, H, D, float(eta), float(magic_num),
                                        int(flags), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, None, "b2p_sched_step")
        return prev, x0

    def _check_ready(self):
        if self.num_inference_steps is None and not self._is_ddpm:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")

    @staticmethod
    def _randn(shape, generator, device, dtype):
        """diffusers.utils.torch_utils.randn_tensor."""
        dev = torch.device(device)
        rdev = torch.device("cpu") if (generator is not None and generator.device.type == "cpu" and dev.type != "cpu") else dev
        return torch.randn(tuple(shape), generator=generator, device=rdev, dtype=dtype).to(dev)


class _GuidanceMixin:
    def _init_guidance(self, cfg):
        self.use_classifier_guidance = (cfg.GUIDANCE.USE_COND == GuidanceType.CLASSIFIER_GUIDANCE.name and cfg.GUIDANCE.LOSS_LIST is not None)
        if self.use_classifier_guidance:
            self.guidance_loss = GuidanceLoss(cfg)

    def _guide(self, model_output, action, target, timestep, eta=0.0):
        if self.use_classifier_guidance and target is not None:
            with torch.enable_grad():
                model_std = torch.tensor(self.coeffs(int(timestep), eta).guidance_grad_scale, device=model_output.device)
                model_output = self.guidance_loss(model_output, action, target, model_std)
        return model_output


class GuidanceDDIMScheduler(_FusedScheduler, _GuidanceMixin):
    """scheduler/guidance_ddim_scheduler.py:13-173."""
    _kind = "guidance_ddim"

    def __init__(self, cfg, **kwargs):
        super().__init__(**kwargs)
        self._init_guidance(cfg)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False, generator=None,
             variance_noise: Optional[torch.Tensor] = None, return_dict: bool = True, target: Optional[torch.Tensor] = None,
             action: Optional[torch.Tensor] = None) -> Union[SchedulerOutput, Tuple]:
        self._check_ready()
        model_output = self._guide(model_output, action, target, timestep, eta)
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                                 " `variance_noise` stays `None`.")
            if variance_noise is None:
                variance_noise = self._randn(model_output.shape, generator, model_output.device, model_output.dtype)
        prev, x0 = self._launch(model_output, timestep, sample, eta=eta, use_clipped=use_clipped_model_output,
                                noise=variance_noise if eta > 0 else None)
        return SchedulerOutput(prev_sample=prev, pred_original_sample=x0) if return_dict else (prev,)


class GuidanceDDPMScheduler(_FusedScheduler, _GuidanceMixin):
    """scheduler/guidance_ddpm_scheduler.py:12-178 (with the missing ``numpy`` import of :41 irrelevant here)."""
    _kind = "guidance_ddpm"
    _is_ddpm = True

    def __init__(self, cfg, **kwargs):
        super().__init__(**kwargs)
        self._init_guidance(cfg)

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True, target: Optional[torch.Tensor] = None,
             action: Optional[torch.Tensor] = None, variance_noise: Optional[torch.Tensor] = None) -> Union[SchedulerOutput, Tuple]:
        """``variance_noise`` is an extension (the reference draws the noise internally with ``generator``)."""
        model_output = self._guide(model_output, action, target, timestep)
        if int(timestep) > 0 and variance_noise is None:
            variance_noise = self._randn(model_output.shape, generator, model_output.device, model_output.dtype)
        prev, x0 = self._launch(model_output, timestep, sample, noise=variance_noise if int(timestep) > 0 else None)
        return SchedulerOutput(prev_sample=prev, pred_original_sample=x0) if return_dict else (prev,)


class InpaintingDDIMScheduler(_FusedScheduler):
    """scheduler/inpainting_ddim_scheduler.py:9-153 (including the scalar-variance offset of :108-128)."""
    _kind = "inpainting_ddim"

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False, generator=None,
             variance_noise: Optional[torch.Tensor] = None, target_traj: Optional[torch.Tensor] = None,
             target_mask: Optional[torch.Tensor] = None, return_dict: bool = True) -> Union[SchedulerOutput, Tuple]:
        self._check_ready()
        blend = target_traj is not None and target_mask is not None
        if eta > 0 and variance_noise is not None and generator is not None:
            raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                             " `variance_noise` stays `None`.")
        noise = variance_noise
        if noise is None and (blend or eta > 0):
            noise = self._randn(model_output.shape, generator, model_output.device, model_output.dtype)
        prev, x0 = self._launch(model_output, timestep, sample, eta=eta, use_clipped=use_clipped_model_output, noise=noise,
                                target_traj=target_traj if blend else None, target_mask=target_mask if blend else None)
        return SchedulerOutput(prev_sample=prev, pred_original_sample=x0) if return_dict else (prev,)


class InpaintingDDPMScheduler(_FusedScheduler):
    """scheduler/inpainting_ddpm_scheduler.py:9-146."""
    _kind = "inpainting_ddpm"
    _is_ddpm = True

    def step(self, model_output, timestep, sample, generator=None, variance_noise: Optional[torch.Tensor] = None,
             target_traj: Optional[torch.Tensor] = None, target_mask: Optional[torch.Tensor] = None,
             return_dict: bool = True) -> Union[SchedulerOutput, Tuple]:
        blend = target_traj is not None and target_mask is not None
        noise = variance_noise
        if noise is None:
            noise = self._randn(model_output.shape, generator, model_output.device, model_output.dtype)
        prev, x0 = self._launch(model_output, timestep, sample, noise=noise, target_traj=target_traj if blend else None,
                                target_mask=target_mask if blend else None)
        return SchedulerOutput(prev_sample=prev, pred_original_sample=x0) if return_dict else (prev,)


SCHEDULER_FUNC = {"ddim": GuidanceDDIMScheduler, "ddpm": GuidanceDDPMScheduler}  # interact.py:48-52 (the dpm entry is CARLA-eval only)
