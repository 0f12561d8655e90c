"""The sampling loop ``interact.py:115-168`` restated for a batch.  TEST INFRASTRUCTURE (oracle/__init__.py).

Batch generalisation: the timestep is repeated to the batch as ``train.py:85`` does; classifier
guidance is the per-sample map of the B=1 rule.  DDPM noise is injected (``noise[i]`` is consumed by loop iteration i).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import guidance as G
from . import schedulers as S
from . import unet as U

MAGIC_NUM = 23.315  # modeling/temporal.py:195


def plan(sd, mode: str, scheduler: str, x_init: torch.Tensor, feat: torch.Tensor, num_inference_steps: int,
         target: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None, free_scale: float = 7.5,
         classifier_scale: float = 15.0, target_traj: Optional[torch.Tensor] = None, target_mask: Optional[torch.Tensor] = None,
         num_train_timesteps: int = 100, postprocess: bool = True, trace: Optional[list] = None,
         sched_overrides: Optional[dict] = None) -> torch.Tensor:
    """scheduler in {guidance_ddim, guidance_ddpm, inpainting_ddim, inpainting_ddpm}; returns trajectories [B,H,D]."""
    cfg = S.SchedCfg(num_train_timesteps=num_train_timesteps, num_inference_steps=num_inference_steps, **(sched_overrides or {}))
    ac = S.alphas_cumprod(num_train_timesteps)
    B = x_init.shape[0]
    trajs = x_init.clone()
    if mode == "FREE_GUIDANCE":
        assert target is not None
        cond = torch.cat([target, torch.zeros_like(target)], dim=0)
    trajs[:, 0, :3] = 0.0
    inpaint = scheduler.startswith("inpainting")
    ddim = scheduler.endswith("ddim")
    for i, t in enumerate(S.leading_timesteps(num_train_timesteps, num_inference_steps)):
        t = int(t)
        tt = torch.full((B,), t, dtype=torch.long)
        action = None
        with torch.no_grad():
            if mode == "FREE_GUIDANCE":
                out = U.unet_forward(sd, torch.cat([trajs, trajs], 0), feat, torch.tensor([t]), cond, mode)
                c, u = out.chunk(2, dim=0)
                mo = u + free_scale * (c - u)
            elif mode == "CLASSIFIER_GUIDANCE":
                action, te = U.unet_forward(sd, trajs, feat, tt, None, mode, return_action_and_time_only=True)
            else:
                mo = U.unet_forward(sd, trajs, feat, tt, None, mode)
        if mode == "CLASSIFIER_GUIDANCE":
            with torch.enable_grad():
                action = action.detach().requires_grad_()
                state = U.traj_predict(sd, action[:, :-1], te)
                state = torch.cat([torch.zeros_like(state[:, :1]), state], dim=1)
                mo = torch.cat([state, action], dim=-1)
                if not inpaint and target is not None:
                    p = t - num_train_timesteps // num_inference_steps
                    var = S.ddim_variance(ac, t, p) if ddim else S.ddpm_variance(ac, t, p)
                    mo = G.guidance_update(mo, action, target, torch.exp(0.5 * var), classifier_scale)  # quirk 4
            mo = mo.detach()
        n = noise[i] if noise is not None else None
        kw = dict(target_traj=target_traj, target_mask=target_mask, inpainting=True) if inpaint else {}
        if ddim:
            trajs, x0 = S.ddim_step(cfg, ac, mo, t, trajs, variance_noise=n, **kw)
        else:
            trajs, x0 = S.ddpm_step(cfg, ac, mo, t, trajs, variance_noise=n, **kw)
        trajs[:, 0, :3] = 0.0
        if trace is not None:
            trace.append(dict(t=t, model_output=mo.clone(), prev_sample=trajs.clone(), x0=x0.clone()))
    if postprocess:
        trajs = trajs.to(torch.float32).clamp(-1, 1)
        trajs[..., :2] *= MAGIC_NUM
    return trajs
