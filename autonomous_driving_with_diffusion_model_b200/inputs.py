"""Caller-side input pipeline.

``preprocess_frames`` replaces ``T.Compose([T.ToTensor(), T.Normalize(mean, std)])`` on the 900x256 camera frame
(interact.py:72-77, 170-172) for uint8 frames that are ALREADY ON THE GPU: uploading the uint8 frame and converting on
the device moves 4x fewer bytes over PCIe than the reference's host-side float conversion.  Same arithmetic in the same
order (u8 -> f32, / 255, - mean, / std), so the result is bit-identical to torchvision's.  The output is the logical
[N,3,H,W] tensor in channels-last memory, which is what the encoder's cuDNN path consumes.

``process_next_waypoint`` is interact.py:185-202 (target point into the ego frame, scaled by the model's magic number)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_frames(frames_u8: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """frames_u8: uint8 CUDA tensor [N,H,W,3] or [H,W,3] (RGB, HWC like the simulator's camera) -> float32 [N,3,H,W]."""
    if frames_u8.device.type != "cuda":
        raise RuntimeError("preprocess_frames needs a CUDA uint8 tensor (no CPU fallback; use torchvision on the host)")
    if frames_u8.dtype != torch.uint8 or frames_u8.shape[-1] != 3 or frames_u8.dim() not in (3, 4):
        raise ValueError("frames must be uint8 [N,H,W,3] or [H,W,3]")
    x = frames_u8.unsqueeze(0) if frames_u8.dim() == 3 else frames_u8
    x = x.contiguous()
    n, h, w, _ = x.shape
    out = torch.empty((n, 3, h, w), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
    if n * h * w == 0:
        return out
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    with torch.cuda.device(x.device):
        rc = _lib.load().b2p_preprocess_frames(_lib.ptr(x), _lib.ptr(out), n * h * w, m3, s3, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, None, "b2p_preprocess_frames")
    return out


def process_next_waypoint(next_point, cur_point, yaw, magic_num: float = 23.315, device=None) -> torch.Tensor:
    """interact.py:185-202: world-frame route point(s) [H,2] -> ego-frame target point(s) [H,2], normalised."""
    if math.isnan(yaw):
        yaw = 0.0
    yaw = yaw + math.pi / 2.0
    R = np.array([[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]])
    local = R.T.dot((np.asarray(next_point) - np.asarray(cur_point)).T).T
    t = torch.FloatTensor(np.stack([local[:, 1] / magic_num, -local[:, 0] / magic_num], axis=-1))
    return t.to(device) if device is not None else t
