"""Closed-form fp32 restatement of the four reference scheduler steps.  TEST INFRASTRUCTURE (oracle/__init__.py).

Sources restated:
  * scheduler/guidance_ddim_scheduler.py:60-173     -> ``ddim_step``      (guidance handled by the caller, see plan.py)
  * scheduler/guidance_ddpm_scheduler.py:59-178     -> ``ddpm_step``
  * scheduler/inpainting_ddim_scheduler.py:10-153   -> ``ddim_step(..., target_traj=, target_mask=)``
  * scheduler/inpainting_ddpm_scheduler.py:10-146   -> ``ddpm_step(..., target_traj=, target_mask=)``
  * diffusers==0.28.0 DDIMScheduler/DDPMScheduler base members (third party, restated):
    betas / alphas_cumprod, ``set_timesteps`` ("leading"), ``_get_variance``, ``previous_timestep``,
    ``_threshold_sample``.

All scalar coefficient arithmetic is done on 0-dim fp32 tensors in the reference's operation order so that the
results are bit-identical to the reference on CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


def alphas_cumprod(num_train_timesteps: int = 100, schedule: str = "squaredcos_cap_v2", beta_start: float = 1e-4,
                   beta_end: float = 0.02) -> torch.Tensor:
    n = num_train_timesteps
    if schedule == "squaredcos_cap_v2":
        bar = lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        betas = torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), 0.999) for i in range(n)], dtype=torch.float32)
    elif schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    elif schedule == "scaled_linear":
        betas = torch.linspace(beta_start**0.5, beta_end**0.5, n, dtype=torch.float32) ** 2
    else:
        raise NotImplementedError(schedule)
    return torch.cumprod(1.0 - betas, dim=0)


def leading_timesteps(num_train_timesteps: int, num_inference_steps: int, steps_offset: int = 0) -> np.ndarray:
    ratio = num_train_timesteps // num_inference_steps
    return (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + steps_offset


def threshold(x0: torch.Tensor, ratio: float = 0.995, sample_max_value: float = 1.0) -> torch.Tensor:
    """Dynamic thresholding over ALL elements of a sample (quirk 8: dim 1 is treated as 'channels')."""
    b = x0.shape[0]
    flat = x0.reshape(b, -1)
    s = torch.quantile(flat.abs(), ratio, dim=1)
    s = torch.clamp(s, min=1, max=sample_max_value).unsqueeze(1)
    return (torch.clamp(flat, -s, s) / s).reshape(x0.shape)


@dataclass
class SchedCfg:
    num_train_timesteps: int = 100
    num_inference_steps: int = 100
    prediction_type: str = "sample"
    thresholding: bool = True
    clip_sample: bool = True
    clip_sample_range: float = 1.0
    dynamic_thresholding_ratio: float = 0.995
    sample_max_value: float = 1.0


def _x0_eps(cfg: SchedCfg, m, x, a_t, need_eps=True):
    b_t = 1 - a_t
    if cfg.prediction_type == "sample":
        x0 = m
        eps = (x - a_t ** 0.5 * x0) / b_t ** 0.5 if need_eps else None  # quirk 2: from the UN-clamped x0
    elif cfg.prediction_type == "epsilon":
        x0 = (x - b_t ** 0.5 * m) / a_t ** 0.5
        eps = m
    elif cfg.prediction_type == "v_prediction":
        x0 = (a_t ** 0.5) * x - (b_t ** 0.5) * m
        eps = (a_t ** 0.5) * m + (b_t ** 0.5) * x
    else:
        raise ValueError(cfg.prediction_type)
    if cfg.thresholding:
        x0 = threshold(x0, cfg.dynamic_thresholding_ratio, cfg.sample_max_value)
    elif cfg.clip_sample:
        x0 = x0.clamp(-cfg.clip_sample_range, cfg.clip_sample_range)
    return x0, eps


def ddim_variance(ac: torch.Tensor, t: int, p: int) -> torch.Tensor:
    a_t = ac[t]
    a_p = ac[p] if p >= 0 else torch.tensor(1.0)
    return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)


def ddpm_variance(ac: torch.Tensor, t: int, p: int) -> torch.Tensor:
    a_t = ac[t]
    a_p = ac[p] if p >= 0 else torch.tensor(1.0)
    return torch.clamp((1 - a_p) / (1 - a_t) * (1 - a_t / a_p), min=1e-20)


def ddim_step(cfg: SchedCfg, ac: torch.Tensor, model_output: torch.Tensor, t: int, sample: torch.Tensor,
              eta: float = 0.0, use_clipped_model_output: bool = False, variance_noise: Optional[torch.Tensor] = None,
              target_traj: Optional[torch.Tensor] = None, target_mask: Optional[torch.Tensor] = None,
              inpainting: bool = False):
    """Returns (prev_sample, pred_original_sample).  ``inpainting=True`` selects the Inpainting class, whose update
    adds the SCALAR variance to the unknown part even without target/mask (quirk 3, inpainting_ddim_scheduler.py:108-128)."""
    t = int(t)
    p = t - cfg.num_train_timesteps // cfg.num_inference_steps
    a_t = ac[t]
    a_p = ac[p] if p >= 0 else torch.tensor(1.0)
    x0, eps = _x0_eps(cfg, model_output, sample, a_t)
    variance = ddim_variance(ac, t, p)
    std = eta * variance ** 0.5
    if use_clipped_model_output:
        eps = (sample - a_t ** 0.5 * x0) / (1 - a_t) ** 0.5
    direction = (1 - a_p - std ** 2) ** 0.5 * eps
    if not inpainting:
        prev = a_p ** 0.5 * x0 + direction
    else:
        unknown = (a_p ** 0.5) * x0 + direction + variance
        if target_traj is not None and target_mask is not None:
            noise = variance_noise
            known = (a_p ** 0.5) * target_traj + ((1.0 - a_p) ** 0.5) * (noise if t > 0 else 0)
            prev = target_mask * known + (1.0 - target_mask) * unknown
        else:
            prev = unknown
    if eta > 0:
        prev = prev + std * variance_noise
    return prev, x0


def ddpm_step(cfg: SchedCfg, ac: torch.Tensor, model_output: torch.Tensor, t: int, sample: torch.Tensor,
              variance_noise: Optional[torch.Tensor] = None, target_traj: Optional[torch.Tensor] = None,
              target_mask: Optional[torch.Tensor] = None, inpainting: bool = False):
    t = int(t)
    p = t - cfg.num_train_timesteps // (cfg.num_inference_steps or cfg.num_train_timesteps)
    a_t = ac[t]
    a_p = ac[p] if p >= 0 else torch.tensor(1.0)
    b_t, b_p = 1 - a_t, 1 - a_p
    cur_a = a_t / a_p
    cur_b = 1 - cur_a
    x0, _ = _x0_eps(cfg, model_output, sample, a_t, need_eps=False)
    c0 = (a_p ** 0.5 * cur_b) / b_t
    c1 = cur_a ** 0.5 * b_p / b_t
    var = ddpm_variance(ac, t, p)
    if not inpainting:
        prev = c0 * x0 + c1 * sample
        noise_term = (var ** 0.5) * variance_noise if t > 0 else 0
        prev = prev + noise_term
    else:
        noise_term = (var ** 0.5) * variance_noise if t > 0 else 0
        unknown = c0 * x0 + c1 * sample + noise_term
        if target_traj is not None and target_mask is not None:
            known = (a_p ** 0.5) * target_traj + ((1.0 - a_p) ** 0.5) * (variance_noise if t > 0 else 0)
            prev = target_mask * known + (1.0 - target_mask) * unknown
        else:
            prev = unknown
    return prev, x0
