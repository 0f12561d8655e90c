"""Deterministic synthetic weights and inputs with the reference's state_dict contract and input shapes (SURVEY.md §8d
"Synthetic inputs"): what bench.py, the scripts, smoke() and the tests feed both the CUDA path and the CPU oracle.
Not a checker and never timed; it lives in the package (not under oracle/) so that nothing on the product side imports
the oracle.

The key names, shapes and registration order restate what the reference constructor produces
(``modeling/temporal.py:59-195``, ``modeling/helpers.py:22-112``, ``modeling/resnet.py:163-296``); the VALUES
come from a counter-based splitmix64 hash so they do not depend on torch's RNG streams or default
initialisers ("never rely on reproducing PyTorch's default initialisers").
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import List, Tuple

import numpy as np
import torch

MODES = ("NO_GUIDANCE", "FREE_GUIDANCE", "CLASSIFIER_GUIDANCE")

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def hash_uniform(tag: str, n: int, seed: int = 0) -> np.ndarray:
    """n float64 values in [0, 1), a pure function of (tag, seed, index)."""
    h = int.from_bytes(hashlib.sha256(f"{tag}/{seed}".encode()).digest()[:8], "little")
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(h)
        bits = _splitmix64(_splitmix64(idx))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def hash_symmetric(tag: str, shape, bound: float, seed: int = 0) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    u = hash_uniform(tag, n, seed)
    return torch.from_numpy(((2.0 * u - 1.0) * bound).astype(np.float32).reshape(shape))


def hash_normal(tag: str, shape, seed: int = 0) -> torch.Tensor:
    """Standard-normal fp32 tensor (Box-Muller on the hash stream)."""
    n = int(np.prod(shape)) if len(shape) else 1
    m = (n + 1) // 2
    u1 = hash_uniform(tag + "/u1", m, seed)
    u2 = hash_uniform(tag + "/u2", m, seed)
    r = np.sqrt(-2.0 * np.log(1.0 - u1))
    z = np.concatenate([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)])[:n]
    return torch.from_numpy(z.astype(np.float32).reshape(shape))


# ----------------------------------------------------------------------------------------------
# state_dict specification: list of (key, shape, kind), kind in
#   "w"  weight matrix / conv kernel  -> uniform(+-1/sqrt(fan_in))
#   "b"  bias                          -> uniform(+-1/sqrt(fan_in of the owning layer))
#   "g"  norm scale                    -> 1 + uniform(+-0.1)
#   "nb" norm bias                     -> uniform(+-0.1)
#   "x"  xavier matrix (TrajPredict)   -> uniform(+-sqrt(6/(fan_in+fan_out)))
#   "he" resnet conv                   -> uniform(+-sqrt(3/fan_in))   (keeps the 16 residual adds from blowing the feature up)
#   "rm"/"rv"/"nt"  BatchNorm running_mean / running_var / num_batches_tracked buffers
# ----------------------------------------------------------------------------------------------
Spec = Tuple[str, Tuple[int, ...], str, int]  # key, shape, kind, fan_in


def _bn(prefix: str, c: int) -> List[Spec]:
    return [
        (f"{prefix}.weight", (c,), "g", 0),
        (f"{prefix}.bias", (c,), "nb", 0),
        (f"{prefix}.running_mean", (c,), "rm", 0),
        (f"{prefix}.running_var", (c,), "rv", 0),
        (f"{prefix}.num_batches_tracked", (), "nt", 0),
    ]


def resnet34_specs(prefix: str = "perception", out_dim: int = 64) -> List[Spec]:
    s: List[Spec] = [(f"{prefix}.conv1.weight", (64, 3, 7, 7), "he", 3 * 49)]
    s += _bn(f"{prefix}.bn1", 64)
    inplanes = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3)), start=1):
        for bi in range(nblk):
            stride = 2 if (bi == 0 and li > 1) else 1
            p = f"{prefix}.layer{li}.{bi}"
            s.append((f"{p}.conv1.weight", (planes, inplanes, 3, 3), "he", inplanes * 9))
            s += _bn(f"{p}.bn1", planes)
            s.append((f"{p}.conv2.weight", (planes, planes, 3, 3), "he", planes * 9))
            s += _bn(f"{p}.bn2", planes)
            if stride != 1 or inplanes != planes:
                s.append((f"{p}.downsample.0.weight", (planes, inplanes, 1, 1), "he", inplanes))
                s += _bn(f"{p}.downsample.1", planes)
            inplanes = planes
    s.append((f"{prefix}.fc.weight", (out_dim, 512), "w", 512))
    s.append((f"{prefix}.fc.bias", (out_dim,), "b", 512))
    return s


def _conv_block(prefix: str, cin: int, cout: int, k: int = 5) -> List[Spec]:
    return [
        (f"{prefix}.block.0.weight", (cout, cin, k), "w", cin * k),
        (f"{prefix}.block.0.bias", (cout,), "b", cin * k),
        (f"{prefix}.block.2.weight", (cout,), "g", 0),
        (f"{prefix}.block.2.bias", (cout,), "nb", 0),
    ]


def _res_block(prefix: str, cin: int, cout: int, embed: int) -> List[Spec]:
    s = _conv_block(f"{prefix}.blocks.0", cin, cout) + _conv_block(f"{prefix}.blocks.1", cout, cout)
    s += [(f"{prefix}.time_mlp.1.weight", (cout, embed), "w", embed), (f"{prefix}.time_mlp.1.bias", (cout,), "b", embed)]
    if cin != cout:
        s += [(f"{prefix}.residual_conv.weight", (cout, cin, 1), "w", cin), (f"{prefix}.residual_conv.bias", (cout,), "b", cin)]
    return s


def traj_predict_specs(prefix: str = "state_pred", hidden: int = 64, in_dim: int = 3, out_dim: int = 4, layers: int = 2) -> List[Spec]:
    s: List[Spec] = [(f"{prefix}.input_proj.weight", (hidden, in_dim), "x", in_dim), (f"{prefix}.input_proj.bias", (hidden,), "b", in_dim)]
    for i in range(layers):
        p = f"{prefix}.encoder_traj.layers.{i}"
        s += [
            (f"{p}.self_attn.in_proj_weight", (3 * hidden, hidden), "x", hidden),
            (f"{p}.self_attn.in_proj_bias", (3 * hidden,), "nb", 0),
            (f"{p}.self_attn.out_proj.weight", (hidden, hidden), "x", hidden),
            (f"{p}.self_attn.out_proj.bias", (hidden,), "nb", 0),
            (f"{p}.linear1.weight", (4 * hidden, hidden), "x", hidden),
            (f"{p}.linear1.bias", (4 * hidden,), "b", hidden),
            (f"{p}.linear2.weight", (hidden, 4 * hidden), "x", 4 * hidden),
            (f"{p}.linear2.bias", (hidden,), "b", 4 * hidden),
            (f"{p}.norm1.weight", (hidden,), "g", 0),
            (f"{p}.norm1.bias", (hidden,), "nb", 0),
            (f"{p}.norm2.weight", (hidden,), "g", 0),
            (f"{p}.norm2.bias", (hidden,), "nb", 0),
        ]
    s += [
        (f"{prefix}.encoder_traj.norm.weight", (hidden,), "g", 0),
        (f"{prefix}.encoder_traj.norm.bias", (hidden,), "nb", 0),
        (f"{prefix}.output_proj.weight", (out_dim, hidden), "x", hidden),
        (f"{prefix}.output_proj.bias", (out_dim,), "b", hidden),
    ]
    return s


def unet_specs(mode: str = "NO_GUIDANCE", transition_dim: int = 7, dim: int = 64, dim_mults=(1, 2, 4, 8), with_perception: bool = True) -> List[Spec]:
    """Registration order: perception, [cond_mlp], time_mlp, downs, ups, mid_block1, mid_block2, head
    (``modeling/temporal.py:83-189``; ``ups`` precedes the mid blocks because both ModuleLists are created at :102-103)."""
    assert mode in MODES
    dims = [transition_dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    embed = 2 * dim
    s: List[Spec] = resnet34_specs("perception", dim) if with_perception else []
    if mode == "FREE_GUIDANCE":
        s += [("cond_mlp.0.weight", (dim, 2), "w", 2), ("cond_mlp.0.bias", (dim,), "b", 2),
              ("cond_mlp.2.weight", (dim, dim), "w", dim), ("cond_mlp.2.bias", (dim,), "b", dim)]
    s += [("time_mlp.1.weight", (4 * dim, dim), "w", dim), ("time_mlp.1.bias", (4 * dim,), "b", dim),
          ("time_mlp.3.weight", (dim, 4 * dim), "w", 4 * dim), ("time_mlp.3.bias", (dim,), "b", 4 * dim)]
    n = len(in_out)
    for i, (ci, co) in enumerate(in_out):
        s += _res_block(f"downs.{i}.0", ci, co, embed) + _res_block(f"downs.{i}.1", co, co, embed)
        if i < n - 1:
            s += [(f"downs.{i}.3.conv.weight", (co, co, 3), "w", co * 3), (f"downs.{i}.3.conv.bias", (co,), "b", co * 3)]
    for i, (ci, co) in enumerate(reversed(in_out[1:])):
        s += _res_block(f"ups.{i}.0", co * 2, ci, embed) + _res_block(f"ups.{i}.1", ci, ci, embed)
        # ConvTranspose1d weight layout is [C_in, C_out, 4] (modeling/helpers.py:89)
        s += [(f"ups.{i}.3.conv.weight", (ci, ci, 4), "w", ci * 4), (f"ups.{i}.3.conv.bias", (ci,), "b", ci * 4)]
    mid = dims[-1]
    s += _res_block("mid_block1", mid, mid, embed) + _res_block("mid_block2", mid, mid, embed)
    fin = in_out[1][0]
    if mode == "CLASSIFIER_GUIDANCE":
        s += _conv_block("act_conv.0", fin, fin) + [("act_conv.1.weight", (3, fin, 1), "w", fin), ("act_conv.1.bias", (3,), "b", fin)]
        s += traj_predict_specs("state_pred", 64, 3, transition_dim - 3, 2)
    else:
        s += _conv_block("final_conv.0", fin, fin) + [("final_conv.1.weight", (transition_dim, fin, 1), "w", fin), ("final_conv.1.bias", (transition_dim,), "b", fin)]
    return s


def make_state_dict(mode: str = "NO_GUIDANCE", seed: int = 0, with_perception: bool = True, **kw) -> "OrderedDict[str, torch.Tensor]":
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape, kind, fan_in in unet_specs(mode, with_perception=with_perception, **kw):
        if kind in ("w", "b"):
            t = hash_symmetric(key, shape, 1.0 / np.sqrt(fan_in), seed)
        elif kind == "he":
            t = hash_symmetric(key, shape, np.sqrt(3.0 / fan_in), seed)
        elif kind == "x":
            t = hash_symmetric(key, shape, np.sqrt(6.0 / (shape[0] + shape[1])), seed)
        elif kind == "g":
            t = 1.0 + hash_symmetric(key, shape, 0.1, seed)
        elif kind == "nb":
            t = hash_symmetric(key, shape, 0.1, seed)
        elif kind == "rm":
            t = hash_symmetric(key, shape, 0.05, seed)
        elif kind == "rv":
            t = 1.0 + hash_symmetric(key, shape, 0.1, seed)
        elif kind == "nt":
            t = torch.tensor(0, dtype=torch.long)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[key] = t
    return sd


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


# ----------------------------------------------------------------------------------------------
# synthetic inputs (SURVEY.md §8d "Synthetic inputs"), also torch-RNG independent
# ----------------------------------------------------------------------------------------------
def synth_inputs(B: int, T: int = 0, seed: int = 0, horizon: int = 16, dim: int = 7, feat_dim: int = 64):
    x = hash_normal("x_init", (B, horizon, dim), seed)
    x[:, 0, :3] = 0.0
    feat = hash_normal("feature", (B, feat_dim), seed)
    target = hash_symmetric("target", (B, 2), 0.5, seed)
    noise = hash_normal("noise", (T, B, horizon, dim), seed) if T else None
    target_traj = hash_symmetric("target_traj", (B, horizon, dim), 1.0, seed)
    mask = torch.zeros(B, horizon, dim)
    mask[:, 0, :3] = 1.0
    mask[:, -1, :2] = 1.0
    return dict(x=x, feat=feat, target=target, noise=noise, target_traj=target_traj, mask=mask)


def synth_image(S: int, seed: int = 0, h: int = 256, w: int = 900) -> torch.Tensor:
    return hash_normal("image", (S, 3, h, w), seed)
