"""The slice of the reference configuration the hot path reads (config.py:9-103 + configs/*.yaml), without yacs.

``load_cfg("configs/guidance/free_guidance.yaml")`` follows the ``_BASE_`` chain like config.py:106-111; with no file it
returns the defaults.  Only keys used by the sampling path are kept."""
from __future__ import annotations

import copy
import os
from types import SimpleNamespace

import yaml

_DEFAULTS = {
    "MODEL": dict(HORIZON=16, TRANSITION_DIM=7, USE_ATTN=False, DIM=64, DIM_MULTS=(1, 2, 4, 8), DIFFUSER_BUILDING_BLOCK="concat"),
    "TRAIN": dict(USE_COND="NO_GUIDANCE", TIME_STEPS=100, SAMPLE_STEPS=100, IMAGE_HEIGHT=256, IMAGE_WIDTH=900,
                  NOISE_SCHEDULER=dict(BETA_START=1e-4, BETA_END=0.02, TYPE="squaredcos_cap_v2", PRED_TYPE="sample")),
    "GUIDANCE": dict(USE_COND="NO_GUIDANCE", LOSS_LIST=None, STEP=1, CLASSIFIER_SCALE=0.1, FREE_SCALE=1.0),
    "EVAL": dict(BATCH_SIZE=4, ETA=0, CHECKPOINT=None, SCHEDULER="ddim", SAMPLE_STEPS=100),
    "PID": dict(TURN_KP=1, TURN_KI=0.5, TURN_KD=1.0, TURN_N=40, SPEED_KP=5, SPEED_KI=0.5, SPEED_KD=1.0, SPEED_N=40),   # config.py:67-76
    "CONTROL": dict(AIM_DIST=4.0, ANGLE_THRESH=0.3, DIST_THRESH=10, BRAKE_SPEED=0.4, BRAKE_RATIO=1.1, CLIP_DELTA=0.25, MAX_THROTTLE=9),   # config.py:79-86
    "B200": dict(PRECISION="fp32", SMALL_BATCH_MAX=4),
}


def _merge(dst: dict, src: dict):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v


def _ns(d):
    return SimpleNamespace(**{k: _ns(v) if isinstance(v, dict) else v for k, v in d.items()})


def _read(path: str) -> dict:
    with open(path) as f:
        d = yaml.safe_load(f) or {}
    base = d.pop("_BASE_", None)
    out = _read(os.path.join(os.path.dirname(path), base)) if base else {}
    _merge(out, d)
    return out


def load_cfg(path: str = None, **overrides):
    d = copy.deepcopy(_DEFAULTS)
    if path:
        _merge(d, _read(path))
    _merge(d, overrides)
    return _ns(d)


def scheduler_kwargs(cfg) -> dict:
    """The keyword set interact.py:81-94 builds."""
    return dict(num_train_timesteps=cfg.TRAIN.SAMPLE_STEPS, prediction_type=cfg.TRAIN.NOISE_SCHEDULER.PRED_TYPE,
                beta_schedule=cfg.TRAIN.NOISE_SCHEDULER.TYPE, beta_start=cfg.TRAIN.NOISE_SCHEDULER.BETA_START,
                beta_end=cfg.TRAIN.NOISE_SCHEDULER.BETA_END, thresholding=True)
