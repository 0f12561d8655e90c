"""Import the REAL reference code from /root/reference.  BUILD-CONTAINER ONLY (the GPU box has no /root/reference).

Used by ``oracle/make_golden.py`` and by the ``not gpu`` tests that validate the restatement (they skip when the
reference tree is absent).  Patches applied:
  * ``modeling.temporal.resnet34(pretrained=True)`` would download ImageNet weights -> forced to ``False``;
  * ``scheduler/guidance_ddpm_scheduler.py:41`` uses ``np`` without importing it -> ``numpy`` injected.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("B2P_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusers_shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modeling", "temporal.py"))


_loaded = None


def load():
    """Returns a namespace with the reference's modules (modeling, scheduler, control, GuidanceType)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("LOCAL_RANK", "1")  # silences the constructor's print (modeling/temporal.py:80)
    import numpy as np
    import modeling.temporal as mt  # noqa: E402
    import modeling.resnet as mr  # noqa: E402

    mt.resnet34 = lambda pretrained=True, **kw: mr.resnet34(pretrained=False, **kw)
    import control  # noqa: E402,F401
    import scheduler  # noqa: E402
    import scheduler.guidance_ddpm_scheduler as gddpm  # noqa: E402

    gddpm.np = np
    from misc.constant import GuidanceType  # noqa: E402

    _loaded = SimpleNamespace(temporal=mt, scheduler=scheduler, control=control, GuidanceType=GuidanceType,
                              gddpm=gddpm)
    return _loaded


def make_cfg(mode: str = "NO_GUIDANCE", classifier_scale: float = 15.0, free_scale: float = 7.5):
    """The slice of the yacs tree the hot path reads (config.py:9-103 + configs/guidance/*.yaml)."""
    g = SimpleNamespace(USE_COND=mode, LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None,
                        STEP=1, CLASSIFIER_SCALE=classifier_scale, FREE_SCALE=free_scale)
    m = SimpleNamespace(HORIZON=16, TRANSITION_DIM=7, USE_ATTN=False, DIM=64, DIM_MULTS=(1, 2, 4, 8),
                        DIFFUSER_BUILDING_BLOCK="concat")
    tr = SimpleNamespace(USE_COND=mode, TIME_STEPS=100, SAMPLE_STEPS=100,
                         NOISE_SCHEDULER=SimpleNamespace(BETA_START=1e-4, BETA_END=0.02, TYPE="squaredcos_cap_v2", PRED_TYPE="sample"))
    return SimpleNamespace(MODEL=m, TRAIN=tr, GUIDANCE=g)


def build_reference_model(mode: str, state_dict):
    ref = load()
    model = ref.temporal.build_model(make_cfg(mode)).eval()
    model.load_state_dict(state_dict, strict=True)
    return model


def build_reference_scheduler(kind: str, mode: str = "NO_GUIDANCE", **over):
    """kind in guidance_ddim / guidance_ddpm / inpainting_ddim / inpainting_ddpm, constructed as interact.py:81-94 does."""
    ref = load()
    cfg = make_cfg(mode)
    kw = dict(num_train_timesteps=100, prediction_type="sample", beta_schedule="squaredcos_cap_v2",
              beta_start=1e-4, beta_end=0.02, thresholding=True)
    kw.update(over)
    cls = {"guidance_ddim": ref.scheduler.GuidanceDDIMScheduler, "guidance_ddpm": ref.scheduler.GuidanceDDPMScheduler,
           "inpainting_ddim": ref.scheduler.InpaintingDDIMScheduler, "inpainting_ddpm": ref.scheduler.InpaintingDDPMScheduler}[kind]
    if kind.startswith("guidance"):
        kw["cfg"] = cfg
    return cls(**kw)
