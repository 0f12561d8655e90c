"""Functional fp32 CPU restatement of the reference denoiser.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Every function takes the reference ``state_dict`` (keys as in SURVEY.md Appendix A) and plain tensors; there are
no nn.Module objects, so this file shares no structure with ``modeling/temporal.py`` — it only has to compute the
same numbers.  Layout is the reference's: trajectories ``[B, H, D]``, activations ``[B, C, L]``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def sinusoidal(t: torch.Tensor, dim: int = 64) -> torch.Tensor:
    """modeling/helpers.py:67-74 — [sin(t*f_i), cos(t*f_i)], f_i = exp(-i*ln(1e4)/(dim/2-1)); always fp32."""
    half = dim // 2
    f = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
    a = t[:, None] * f[None, :]
    return torch.cat((a.sin(), a.cos()), dim=-1)


def conv_gn_mish(sd: SD, p: str, x: torch.Tensor, groups: int = 8) -> torch.Tensor:
    """modeling/helpers.py:95-112 — Conv1d(k, pad k//2) -> GroupNorm(8) over (C/8 x L) per sample -> Mish."""
    w = sd[f"{p}.block.0.weight"]
    y = F.conv1d(x, w, sd[f"{p}.block.0.bias"], padding=w.shape[-1] // 2)
    y = F.group_norm(y, groups, sd[f"{p}.block.2.weight"], sd[f"{p}.block.2.bias"], eps=1e-5)
    return F.mish(y)


def residual_block(sd: SD, p: str, x: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    """modeling/temporal.py:46-55."""
    temb = F.linear(F.mish(cond), sd[f"{p}.time_mlp.1.weight"], sd[f"{p}.time_mlp.1.bias"])
    h = conv_gn_mish(sd, f"{p}.blocks.0", x) + temb[:, :, None]
    h = conv_gn_mish(sd, f"{p}.blocks.1", h)
    if f"{p}.residual_conv.weight" in sd:
        x = F.conv1d(x, sd[f"{p}.residual_conv.weight"], sd[f"{p}.residual_conv.bias"])
    return h + x


def time_embedding(sd: SD, t: torch.Tensor) -> torch.Tensor:
    """modeling/temporal.py:93-98."""
    e = sinusoidal(t.to(torch.float32) if not t.is_floating_point() else t, sd["time_mlp.1.weight"].shape[1])
    e = F.mish(F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def traj_predict(sd: SD, action: torch.Tensor, time_embed: torch.Tensor, p: str = "state_pred", heads: int = 4) -> torch.Tensor:
    """modeling/helpers.py:22-59 — 2-layer post-LN transformer encoder (SiLU FFN, eval => no dropout), final LN, Linear."""
    B, S, _ = action.shape
    d = sd[f"{p}.input_proj.weight"].shape[0]
    pos = sinusoidal(torch.arange(S, device=action.device).float(), d)
    x = F.linear(action, sd[f"{p}.input_proj.weight"], sd[f"{p}.input_proj.bias"]) + pos[None] + time_embed[:, None, :]
    li = 0
    while f"{p}.encoder_traj.layers.{li}.linear1.weight" in sd:
        q = f"{p}.encoder_traj.layers.{li}"
        qkv = F.linear(x, sd[f"{q}.self_attn.in_proj_weight"], sd[f"{q}.self_attn.in_proj_bias"])
        qh, kh, vh = (z.reshape(B, S, heads, d // heads).transpose(1, 2) for z in qkv.chunk(3, dim=-1))
        att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(d // heads), dim=-1) @ vh
        att = att.transpose(1, 2).reshape(B, S, d)
        x = F.layer_norm(x + F.linear(att, sd[f"{q}.self_attn.out_proj.weight"], sd[f"{q}.self_attn.out_proj.bias"]),
                         (d,), sd[f"{q}.norm1.weight"], sd[f"{q}.norm1.bias"], 1e-5)
        ff = F.linear(F.silu(F.linear(x, sd[f"{q}.linear1.weight"], sd[f"{q}.linear1.bias"])), sd[f"{q}.linear2.weight"], sd[f"{q}.linear2.bias"])
        x = F.layer_norm(x + ff, (d,), sd[f"{q}.norm2.weight"], sd[f"{q}.norm2.bias"], 1e-5)
        li += 1
    x = F.layer_norm(x, (d,), sd[f"{p}.encoder_traj.norm.weight"], sd[f"{p}.encoder_traj.norm.bias"], 1e-5)
    return F.linear(x, sd[f"{p}.output_proj.weight"], sd[f"{p}.output_proj.bias"])


def resnet34_feature(sd: SD, img: torch.Tensor, p: str = "perception") -> torch.Tensor:
    """modeling/resnet.py:277-293 with BatchNorm in eval mode; fc replaced by Linear(512->64) (temporal.py:84)."""

    def bn(x, q):
        return F.batch_norm(x, sd[f"{q}.running_mean"], sd[f"{q}.running_var"], sd[f"{q}.weight"], sd[f"{q}.bias"], False, 0.0, 1e-5)

    x = F.relu(bn(F.conv2d(img, sd[f"{p}.conv1.weight"], None, 2, 3), f"{p}.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, nblk in enumerate((3, 4, 6, 3), start=1):
        for bi in range(nblk):
            q = f"{p}.layer{li}.{bi}"
            stride = 2 if (bi == 0 and li > 1) else 1
            y = F.relu(bn(F.conv2d(x, sd[f"{q}.conv1.weight"], None, stride, 1), f"{q}.bn1"))
            y = bn(F.conv2d(y, sd[f"{q}.conv2.weight"], None, 1, 1), f"{q}.bn2")
            if f"{q}.downsample.0.weight" in sd:
                x = bn(F.conv2d(x, sd[f"{q}.downsample.0.weight"], None, stride, 0), f"{q}.downsample.1")
            x = F.relu(y + x)
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)
    return F.linear(x, sd[f"{p}.fc.weight"], sd[f"{p}.fc.bias"])


def unet_forward(sd: SD, x: torch.Tensor, feat: torch.Tensor, t: torch.Tensor, cond: Optional[torch.Tensor] = None,
                 mode: str = "NO_GUIDANCE", return_action_and_time_only: bool = False):
    """modeling/temporal.py:197-245 with the image feature ``feat`` [B or S, 64] already computed
    (hoisting ``self.perception(img)`` is result-identical in eval mode)."""
    h = x.transpose(1, 2)  # b h t -> b t h
    te = time_embedding(sd, t)
    if mode == "FREE_GUIDANCE":
        cond = cond if cond is not None else torch.zeros((x.shape[0], 2), device=x.device)
        if te.shape[0] != cond.shape[0]:
            te = te.repeat(cond.shape[0] // te.shape[0], 1)
        if feat.shape[0] != cond.shape[0]:
            feat = feat.repeat(cond.shape[0] // feat.shape[0], 1)
        c = F.mish(F.linear(cond, sd["cond_mlp.0.weight"], sd["cond_mlp.0.bias"]))
        te = te + F.linear(c, sd["cond_mlp.2.weight"], sd["cond_mlp.2.bias"])
    ci = torch.cat([te, feat], dim=-1)
    skips = []
    i = 0
    while f"downs.{i}.0.blocks.0.block.0.weight" in sd:
        h = residual_block(sd, f"downs.{i}.0", h, ci)
        h = residual_block(sd, f"downs.{i}.1", h, ci)
        skips.append(h)
        if f"downs.{i}.3.conv.weight" in sd:
            h = F.conv1d(h, sd[f"downs.{i}.3.conv.weight"], sd[f"downs.{i}.3.conv.bias"], stride=2, padding=1)
        i += 1
    h = residual_block(sd, "mid_block1", h, ci)
    h = residual_block(sd, "mid_block2", h, ci)
    i = 0
    while f"ups.{i}.0.blocks.0.block.0.weight" in sd:
        h = torch.cat((h, skips.pop()), dim=1)
        h = residual_block(sd, f"ups.{i}.0", h, ci)
        h = residual_block(sd, f"ups.{i}.1", h, ci)
        h = F.conv_transpose1d(h, sd[f"ups.{i}.3.conv.weight"], sd[f"ups.{i}.3.conv.bias"], stride=2, padding=1)
        i += 1
    if mode == "CLASSIFIER_GUIDANCE":
        a = conv_gn_mish(sd, "act_conv.0", h)
        action = F.conv1d(a, sd["act_conv.1.weight"], sd["act_conv.1.bias"]).transpose(1, 2)
        if return_action_and_time_only:
            return action, te
        state = traj_predict(sd, action.detach()[:, :-1], te)
        state = torch.cat([torch.zeros_like(state[:, :1]), state], dim=1)
        return torch.cat([state, action], dim=-1)
    y = conv_gn_mish(sd, "final_conv.0", h)
    return F.conv1d(y, sd["final_conv.1.weight"], sd["final_conv.1.bias"]).transpose(1, 2)
