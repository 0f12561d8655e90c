"""Restatement of classifier guidance.  TEST INFRASTRUCTURE (oracle/__init__.py).

control/guidance_loss.py:10-22 (TargetGuidance) is only defined for B=1 and one target point (it raises for B>1:
"Boolean value of Tensor with more than one value is ambiguous"); the batched semantics used here and by the CUDA
path is the per-sample map of that B=1 rule.  control/guidance.py:35-59 (GuidanceLoss.forward,
GUIDANCE.STEP == 1).
"""
from __future__ import annotations

import torch


def choose_index(x_row: torch.Tensor, target: torch.Tensor) -> int:
    """x_row [H, >=2], target [2] -> waypoint index whose squared distance to the target is penalised."""
    xy = x_row[:, :2]
    target_to_agent = torch.norm(target - xy[0])
    final_to_agent = torch.norm(xy[-1] - xy[0])
    if final_to_agent < target_to_agent:
        return 0  # "dummy point to prevent erratic update" (guidance_loss.py:18-19)
    return int(((xy - target[None]) ** 2).sum(-1).argmin())


def target_loss(x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Sum over the batch of ||x[b, idx_b, :2] - target_b||^2."""
    total = x.new_zeros(())
    for b in range(x.shape[0]):
        idx = choose_index(x[b].detach(), target[b].detach())
        total = total + ((x[b, idx, :2] - target[b]) ** 2).sum()
    return total


def guidance_update(x_guidance: torch.Tensor, action: torch.Tensor, target: torch.Tensor, grad_scale, scale: float) -> torch.Tensor:
    """x_guidance = cat[state(action), action] with a live autograd graph to ``action`` (interact.py:154-160)."""
    with torch.enable_grad():
        loss = target_loss(x_guidance, target)
        state_grad, action_grad = torch.autograd.grad([loss], [x_guidance, action], allow_unused=True)
    if action_grad is None:
        action_grad = torch.zeros_like(action)
    grad = torch.cat([state_grad[..., :-3], action_grad], dim=-1)
    if grad_scale is not None:
        grad = grad * grad_scale
    x = x_guidance.detach().clone()
    x[..., :-3] = x[..., :-3] - scale / 15 * grad[..., :-3]   # quirk 5: state step uses scale/15
    x[..., -3:] = x[..., -3:] - scale * grad[..., -3:]
    return x.clip(-1, 1)
