"""Pieces shared by the two shimmed base classes (restated from diffusers 0.28.0)."""
import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return (self.prev_sample, self.pred_original_sample)[i]


def cosine_betas(n: int, max_beta: float = 0.999) -> torch.Tensor:
    bar = lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
    return torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), max_beta) for i in range(n)], dtype=torch.float32)


def make_betas(schedule: str, n: int, beta_start: float, beta_end: float, trained_betas=None) -> torch.Tensor:
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=torch.float32)
    if schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(beta_start**0.5, beta_end**0.5, n, dtype=torch.float32) ** 2
    if schedule == "squaredcos_cap_v2":
        return cosine_betas(n)
    raise NotImplementedError(f"{schedule} is not implemented")


class SchedulerBase:
    """Config registration + the members both DDIM and DDPM bases share."""

    def _register(self, **kw):
        self.config = SimpleNamespace(**kw)

    def _init_tables(self):
        c = self.config
        self.betas = make_betas(c.beta_schedule, c.num_train_timesteps, c.beta_start, c.beta_end, c.trained_betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, c.num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _leading_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {c.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {c.num_train_timesteps} timesteps."
            )
        self.num_inference_steps = num_inference_steps
        ratio = c.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + c.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def _threshold_sample(self, sample: torch.Tensor) -> torch.Tensor:
        dtype = sample.dtype
        batch_size, channels, *remaining = sample.shape
        if dtype not in (torch.float32, torch.float64):
            sample = sample.float()
        sample = sample.reshape(batch_size, channels * int(np.prod(remaining)))
        s = torch.quantile(sample.abs(), self.config.dynamic_thresholding_ratio, dim=1)
        s = torch.clamp(s, min=1, max=self.config.sample_max_value).unsqueeze(1)
        sample = torch.clamp(sample, -s, s) / s
        return sample.reshape(batch_size, channels, *remaining).to(dtype)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        a = ac[timesteps] ** 0.5
        b = (1 - ac[timesteps]) ** 0.5
        while a.dim() < original_samples.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise
