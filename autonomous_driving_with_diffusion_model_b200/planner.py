"""Whole-plan fast path: the reference's ``Agent.generate_traj`` (interact.py:115-168, identical at
e2e_driving/diffusion_agent.py:179-232) as ONE C-ABI call that replays a captured CUDA graph of the T-step loop
(denoiser + CFG mix / classifier guidance + scheduler step + waypoint overwrite + final clamp/scale)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .constant import GuidanceType
from .modeling import TemporalMapUnet


class DiffusionPlanner:
    """``planner.generate_traj(image, target)`` == the reference agent's method; ``plan`` is the batched form.

    ``scheduler`` is one of this package's scheduler objects (its config + kind select the fused step); ``cfg`` supplies
    EVAL.SAMPLE_STEPS and GUIDANCE.{FREE_SCALE, CLASSIFIER_SCALE}.
    """

    def __init__(self, model: TemporalMapUnet, scheduler, cfg=None, num_inference_steps: Optional[int] = None,
                 free_scale: Optional[float] = None, classifier_scale: Optional[float] = None, use_graph: bool = True):
        self.model, self.scheduler = model, scheduler
        self.num_inference_steps = num_inference_steps if num_inference_steps is not None else cfg.EVAL.SAMPLE_STEPS
        self.free_scale = free_scale if free_scale is not None else (cfg.GUIDANCE.FREE_SCALE if cfg is not None else 1.0)
        self.classifier_scale = classifier_scale if classifier_scale is not None else (cfg.GUIDANCE.CLASSIFIER_SCALE if cfg is not None else 0.1)
        self.use_graph = use_graph
        self.init_trajs = None  # set lazily like interact.py:95-100 (one fixed noise draw per agent)

    def plan_config(self, postprocess: bool = True, eta: float = 0.0) -> "_lib.PlanConfig":
        pc = _lib.PlanConfig()
        pc.sched = self.scheduler.sched_config()
        pc.num_inference_steps = int(self.num_inference_steps)
        pc.eta, pc.free_scale, pc.classifier_scale = float(eta), float(self.free_scale), float(self.classifier_scale)
        pc.magic_num, pc.postprocess, pc.use_graph = float(self.model.magic_num), int(postprocess), int(self.use_graph)
        return pc

    def _needs_noise(self, has_blend: bool) -> bool:
        kind = self.scheduler._kind
        return kind.endswith("ddpm") or (kind.startswith("inpainting") and has_blend)

    @torch.no_grad()
    def plan(self, x_init: torch.Tensor, image_or_feature: torch.Tensor, target: Optional[torch.Tensor] = None,
             noise: Optional[torch.Tensor] = None, target_traj: Optional[torch.Tensor] = None, target_mask: Optional[torch.Tensor] = None,
             generator=None, postprocess: bool = True) -> torch.Tensor:
        """x_init [B,H,D] initial noise; image [S,3,h,w] or feature [S,dim] with S in {1,B}; target [B,2] (CFG /
        classifier); noise [T,B,H,D] for DDPM / inpainting (drawn with ``generator`` when omitted).  Returns [B,H,D]."""
        m = self.model
        if x_init.device.type != "cuda":
            raise RuntimeError("DiffusionPlanner.plan needs CUDA tensors (no CPU fallback)")
        dev = x_init.device
        B, T = x_init.shape[0], int(self.num_inference_steps)
        if x_init.dim() != 3 or tuple(x_init.shape[1:]) != (m.horizon, m.transition_dim):
            raise ValueError(f"x_init must be [B,{m.horizon},{m.transition_dim}], got {tuple(x_init.shape)}")
        if B == 0:
            return x_init.new_zeros((0, m.horizon, m.transition_dim), dtype=torch.float32)
        h = m._handle_for(dev)
        f32 = lambda t, shape=None: None if t is None else (t.detach().to(dev, torch.float32).expand(*shape) if shape else t.detach().to(dev, torch.float32)).contiguous()  # noqa: E731
        feat = f32(m.encode(image_or_feature.to(dev)), (B, m.dim))
        blend = target_traj is not None and target_mask is not None
        if m.use_cond == GuidanceType.NO_GUIDANCE:
            target = None
        if noise is None and self._needs_noise(blend):
            noise = self.scheduler._randn((T, B, m.horizon, m.transition_dim), generator, dev, torch.float32)
        hd = (B, m.horizon, m.transition_dim)
        x, tg, nz = f32(x_init), f32(target, (B, 2)), f32(noise)
        tj, mk = (f32(target_traj, hd), f32(target_mask, hd)) if blend else (None, None)
        out = torch.empty_like(x)
        pc = self.plan_config(postprocess)
        rc = _lib.load().b2p_plan(h, C.byref(pc), _lib.ptr(x), _lib.ptr(feat), _lib.ptr(tg), _lib.ptr(nz), _lib.ptr(tj), _lib.ptr(mk),
                                  _lib.ptr(out), B, m._stream())
        _lib.check(rc, h, "b2p_plan")
        return out

    def plan_host(self, x_init: torch.Tensor, feature: torch.Tensor, target=None, noise=None, target_traj=None, target_mask=None,
                  out: Optional[torch.Tensor] = None, device=None, postprocess: bool = True) -> torch.Tensor:
        """End-to-end entry on HOST tensors (pinned or pageable): H2D, the captured loop, D2H, synchronised on return
        (``b2p_plan_host``).  This is what a non-PyTorch caller of the C ABI would use."""
        m = self.model
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        h = m._handle_for(dev)
        B = x_init.shape[0]
        for t in (x_init, feature, target, noise, target_traj, target_mask):
            if t is not None and (t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous()):
                raise ValueError("plan_host takes contiguous fp32 CPU tensors")
        if out is None:
            out = torch.empty_like(x_init)
        pc = self.plan_config(postprocess)
        with torch.cuda.device(dev):
            rc = _lib.load().b2p_plan_host(h, C.byref(pc), _lib.ptr(x_init), _lib.ptr(feature), _lib.ptr(target), _lib.ptr(noise),
                                           _lib.ptr(target_traj), _lib.ptr(target_mask), _lib.ptr(out), B)
        _lib.check(rc, h, "b2p_plan_host")
        return out

    def generate_traj(self, image: torch.Tensor, target: Optional[torch.Tensor] = None) -> torch.Tensor:
        """interact.py:115-168: one trajectory from the agent's fixed initial noise."""
        if self.model.training:     # Module.eval() walks ~350 submodules (0.5 ms): only when it changes something
            self.model.eval()
        dev = next(self.model.parameters()).device
        if self.init_trajs is None:
            self.init_trajs = torch.randn((1, self.model.horizon, self.model.transition_dim), device=dev)
        if target is not None:
            target = target.reshape(-1, 2)
        return self.plan(self.init_trajs.clone(), image.to(dev), target=target)

    def last_launch_count(self) -> int:
        return self.model.last_launch_count()
