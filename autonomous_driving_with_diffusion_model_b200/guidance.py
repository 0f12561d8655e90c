"""Classifier guidance behind the reference's ``GuidanceLoss`` surface (control/guidance.py:18-59).

This module is the drop-in for the UNMODIFIED loop (interact.py:154-163): the same statements as the reference, with
torch.autograd running through our ``state_pred`` (its backward is the hand-written VJP kernel ``b2p_state_pred_vjp``).
The fully fused form — TargetGuidance index rule, analytic gradient through TrajPredict, scaled update and clip in ONE
kernel — is what ``DiffusionPlanner.plan`` / ``b2p_plan`` run inside the captured loop, and is exported for non-Python
hosts as ``b2p_classifier_guidance``.
"""
from __future__ import annotations

from typing import Union

import torch
import torch.nn as nn


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

    def forward(self, x_guidance: torch.Tensor, action: torch.Tensor, target: torch.Tensor,
                grad_scale: Union[float, torch.Tensor] = None) -> torch.Tensor:
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
