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
from .sharding import shard_bounds


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
        self._stage = {}        # pinned host staging buffers of plan_sharded, keyed by (name, shape)
        self._seeded = set()    # handles whose in-kernel noise stream has been seeded from torch's generator

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
        elif m.use_cond == GuidanceType.CLASSIFIER_GUIDANCE:
            # the reference only guides when the SCHEDULER was built for it (guidance_ddim_scheduler.py:19-21, 89-92):
            # GUIDANCE.USE_COND == CLASSIFIER_GUIDANCE and a LOSS_LIST; inpainting schedulers never guide
            if not getattr(self.scheduler, "use_classifier_guidance", False):
                target = None
            elif target is not None and int(self.scheduler.guidance_loss.guidance_step) != 1:
                raise NotImplementedError("GUIDANCE.STEP != 1: the reference's second guidance iteration differentiates through an "
                                          "`action` that is no longer in the graph (control/guidance.py:42-48) and fails")
        if noise is None and self._needs_noise(blend):
            if generator is not None:   # the reference's randn_tensor(generator=...) stream (guidance_ddpm_scheduler.py:154-157)
                noise = self.scheduler._randn((T, B, m.horizon, m.transition_dim), generator, dev, torch.float32)
            else:                       # no [T,B,H,D] tensor at all: the scheduler kernel draws its own noise (Philox) inside the graph
                self._seed_device_noise(h)
        hd = (B, m.horizon, m.transition_dim)
        x, tg, nz = f32(x_init), f32(target, (B, 2)), f32(noise)
        tj, mk = (f32(target_traj, hd), f32(target_mask, hd)) if blend else (None, None)
        out = torch.empty_like(x)
        pc = self.plan_config(postprocess)
        rc = _lib.load().b2p_plan(h, C.byref(pc), _lib.ptr(x), _lib.ptr(feat), _lib.ptr(tg), _lib.ptr(nz), _lib.ptr(tj), _lib.ptr(mk),
                                  _lib.ptr(out), B, m._stream())
        _lib.check(rc, h, "b2p_plan")
        return out

    def _seed_device_noise(self, h) -> None:
        """First use on a handle: key the in-kernel Philox stream from torch's default CPU generator, so that
        ``torch.manual_seed(s)`` followed by the same call sequence reproduces the same plans."""
        if h.value not in self._seeded:
            self.seed_noise(int(torch.randint(0, 2 ** 62, (1,)).item()), handle=h)

    def seed_noise(self, seed: int, handle=None) -> None:
        """Seed the in-kernel noise of every device handle (or one) and reset its per-plan counter."""
        lib = _lib.load()
        for h in ([handle] if handle is not None else list(self.model._handles.values())):
            _lib.check(lib.b2p_set_noise_seed(h, C.c_uint64(seed & (2 ** 64 - 1))), h, "b2p_set_noise_seed")
            self._seeded.add(h.value)

    def _pinned(self, name: str, src: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if src is None:
            return None
        if src.device.type != "cpu" or src.dtype != torch.float32:
            raise ValueError("plan_sharded takes fp32 CPU tensors")
        src = src.contiguous()
        if src.is_pinned():
            return src
        key = (name, tuple(src.shape))
        buf = self._stage.get(key)
        if buf is None:
            for k in [k for k in self._stage if k[0] == name]:
                del self._stage[k]
            buf = self._stage[key] = torch.empty(src.shape, dtype=torch.float32).pin_memory()
        buf.copy_(src)
        return buf

    def plan_sharded(self, x_init: torch.Tensor, feature: torch.Tensor, target=None, noise=None, target_traj=None, target_mask=None,
                     devices=None, out: Optional[torch.Tensor] = None, postprocess: bool = True) -> torch.Tensor:
        """ONE host batch planned on several GPUs of one box from ONE process (north_star "independent planning
        requests are sharded by batch across the 8 GPUs of one box with no NCCL on the sampling path"): contiguous split
        (``sharding.shard_bounds``), one C handle + private stream + captured graph per device, inputs staged through pinned
        host memory, all shards enqueued back to back (``b2p_plan_sharded_host``), joined, results concatenated in the
        pinned output.  Host fp32 tensors in, host tensor out.  No operation of the path crosses samples, so the result is
        bitwise the single-GPU result as long as every shard runs the same kernels as the whole batch would: not the
        small-batch GEMV program (<= ``small_batch_max`` trajectories) against the tile kernels, and the same column-tile
        widths (the tile kernels widen their tiles once a layer exceeds one wave of CTAs, beyond ~256 trajectories; plans of
        different tile widths agree to bf16x3 rounding, not bitwise)."""
        m = self.model
        devs = list(range(torch.cuda.device_count())) if devices is None else [torch.device(d).index if not isinstance(d, int) else d for d in devices]
        if not devs:
            raise RuntimeError("plan_sharded needs at least one CUDA device (no CPU fallback)")
        B = x_init.shape[0]
        if tuple(x_init.shape[1:]) != (m.horizon, m.transition_dim):
            raise ValueError(f"x_init must be [B,{m.horizon},{m.transition_dim}], got {tuple(x_init.shape)}")
        if feature.dim() != 2 or feature.shape[0] != B:
            raise ValueError("plan_sharded takes the precomputed [B, dim] feature (encode scenes once per device batch beforehand)")
        if m.use_cond == GuidanceType.NO_GUIDANCE:
            target = None
        elif m.use_cond == GuidanceType.CLASSIFIER_GUIDANCE and not getattr(self.scheduler, "use_classifier_guidance", False):
            target = None
        blend = target_traj is not None and target_mask is not None
        if B == 0:
            return x_init.new_zeros((0, m.horizon, m.transition_dim))
        handles = (C.c_void_p * len(devs))(*[m._handle_for(torch.device("cuda", d)) for d in devs])
        if noise is None and self._needs_noise(blend):
            for h in handles:
                self._seed_device_noise(C.c_void_p(h))
        x, f, tg, nz = self._pinned("x", x_init), self._pinned("feat", feature), self._pinned("target", target), self._pinned("noise", noise)
        tj, mk = (self._pinned("traj", target_traj), self._pinned("mask", target_mask)) if blend else (None, None)
        if out is None or not out.is_pinned() or tuple(out.shape) != tuple(x.shape):
            res = self._stage.get(("out", tuple(x.shape)))
            if res is None:
                res = self._stage[("out", tuple(x.shape))] = torch.empty(x.shape, dtype=torch.float32).pin_memory()
        else:
            res = out
        pc = self.plan_config(postprocess)
        rc = _lib.load().b2p_plan_sharded_host(handles, len(devs), C.byref(pc), _lib.ptr(x), _lib.ptr(f), _lib.ptr(tg), _lib.ptr(nz),
                                               _lib.ptr(tj), _lib.ptr(mk), _lib.ptr(res), B)
        _lib.check(rc, C.c_void_p(handles[0]), "b2p_plan_sharded_host")
        if out is not None and out is not res:
            out.copy_(res)
            return out
        return res if out is res else res.clone()

    def shard_sizes(self, batch: int, devices=None):
        """Trajectories per device for ``plan_sharded`` (contiguous split, first ``batch % n`` devices take one more)."""
        n = torch.cuda.device_count() if devices is None else len(devices)
        return [hi - lo for lo, hi in shard_bounds(batch, n)]

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
