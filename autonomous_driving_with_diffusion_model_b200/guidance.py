"""Classifier guidance behind the reference's ``GuidanceLoss`` surface (control/guidance.py:18-59).

Fast path: when the guided tensor was assembled the way interact.py:154-160 does (``cat[cat[0, state_pred(action[:, :-1],
time_embed)], action]``) the whole update — TargetGuidance index rule, analytic gradient through TrajPredict, scaled
update, clip — is ONE kernel (``b2p_classifier_guidance``).  Otherwise the generic path below reproduces the reference
with torch.autograd (our ``state_pred`` is autograd-aware through its VJP kernel).
"""
from __future__ import annotations

import ctypes as C
from typing import Union

import torch
import torch.nn as nn

from . import _lib


def convert(loss_config):
    it = iter(loss_config)
    return dict(zip(it, it))


class TargetGuidance(nn.Module):
    """control/guidance_loss.py:5-22, generalised to a batch as the per-sample map of the B=1 rule."""

    def forward(self, x, target):
        total = x.new_zeros(())
        for b in range(x.shape[0]):
            xy, tg = x[b, :, :2], target.reshape(-1, 2)[b if target.reshape(-1, 2).shape[0] > 1 else 0]
            if torch.norm(xy[-1] - xy[0]) < torch.norm(tg - xy[0]):
                idx = 0
            else:
                idx = int(((xy.detach() - tg) ** 2).sum(-1).argmin())
            total = total + ((xy[idx] - tg) ** 2).sum()
        return total


_LOSSES = {"TargetGuidance": TargetGuidance}


class GuidanceLoss(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.loss_list = nn.ModuleList([_LOSSES[name](**convert(conf)) for name, conf in cfg.GUIDANCE.LOSS_LIST])
        self.guidance_step = cfg.GUIDANCE.STEP
        self.scale = cfg.GUIDANCE.CLASSIFIER_SCALE

    def compute_loss(self, x, target):
        total = 0
        for loss in self.loss_list:
            total = total + loss(x, target)
        return total

    def _fused_context(self, x_guidance, action):
        """Returns (model, time_embed) if ``x_guidance`` carries the provenance the planner / model attach, else None."""
        ctx = getattr(x_guidance, "_b2p_guidance_ctx", None)
        if ctx is None or self.guidance_step != 1 or len(self.loss_list) != 1 or not isinstance(self.loss_list[0], TargetGuidance):
            return None
        return ctx

    def forward(self, x_guidance: torch.Tensor, action: torch.Tensor, target: torch.Tensor,
                grad_scale: Union[float, torch.Tensor] = None) -> torch.Tensor:
        ctx = self._fused_context(x_guidance, action)
        if ctx is not None:
            model, time_embed = ctx
            x = x_guidance.detach().contiguous().float()
            B = x.shape[0]
            tg = target.detach().to(x.device, torch.float32).reshape(-1, 2).expand(B, 2).contiguous()
            gs = 1.0 if grad_scale is None else float(grad_scale)
            h = model._handle_for(x.device)
            rc = _lib.load().b2p_classifier_guidance(h, _lib.ptr(x), _lib.ptr(time_embed), _lib.ptr(tg), gs, float(self.scale), B,
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, h, "b2p_classifier_guidance")
            return x
        # generic path: same statements as the reference, autograd through our VJP-backed state_pred
        for _ in range(self.guidance_step):
            with torch.enable_grad():
                if not x_guidance.requires_grad:
                    x_guidance.requires_grad_()
                loss = self.compute_loss(x_guidance, target)
                state_grad, action_grad = torch.autograd.grad([loss], [x_guidance, action], allow_unused=True)
                if action_grad is None:
                    action_grad = torch.zeros_like(action)
                grad = torch.cat([state_grad[..., :-3], action_grad], dim=-1)
            if grad_scale is not None:
                grad = grad * grad_scale
            x_guidance = x_guidance.detach().clone()
            x_guidance[..., :-3] = x_guidance[..., :-3] - self.scale / 15 * grad[..., :-3]
            x_guidance[..., -3:] = x_guidance[..., -3:] - self.scale * grad[..., -3:]
        x_guidance.requires_grad_(False)
        return x_guidance.clip(-1, 1)
