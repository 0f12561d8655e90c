"""Checkpoint ingestion in the reference's format.

The reference saves ``{"state_dict", "optimizer", "lr_scheduler", "iter", "ema_state_dict"}`` (train.py:288-298) and, at
inference, loads ``state_dict`` and then overwrites every parameter POSITIONALLY with the EMA shadow parameters
(interact.py:102-108, misc/load_param.py:4-8).  The positional copy only works because ``model.parameters()`` enumerates
in the reference constructor's registration order — tests/test_abi.py pins that order against the real reference."""
from __future__ import annotations

from typing import Iterable, Mapping, Union

import torch


def copy_parameters(from_parameters: Iterable[torch.Tensor], to_parameters: Iterable[torch.nn.Parameter]) -> None:
    """misc/load_param.py:4-8 (same name, same argument meaning, same length assertion) plus a shape check."""
    src, dst = list(from_parameters), list(to_parameters)
    assert len(src) == len(dst), f"EMA shadow parameter count {len(src)} != model parameter count {len(dst)}"
    with torch.no_grad():
        for i, (s, p) in enumerate(zip(src, dst)):
            if tuple(s.shape) != tuple(p.shape):
                raise ValueError(f"EMA shadow parameter {i} has shape {tuple(s.shape)}, the model expects {tuple(p.shape)}: "
                                 "the checkpoint was trained with a different configuration")
            p.copy_(s.detach().to(p.device))    # NOT p.data.copy_: a write through .data does not bump the version counter
    owners = {id(o): o for o in (getattr(p, "_b2p_owner", lambda: None)() for p in dst) if o is not None}
    for o in owners.values():                   # packed device copies / folded encoder weights are rebuilt on the next call
        o.invalidate_weights()


def load_checkpoint(model: torch.nn.Module, checkpoint: Union[str, Mapping], use_ema: bool = True, map_location="cpu") -> dict:
    """interact.py:102-108: ``model.load_state_dict(weight["state_dict"])`` then the positional EMA copy.  ``checkpoint`` is a
    path (``cfg.EVAL.CHECKPOINT``) or an already loaded dict.  Returns the non-tensor metadata (``iter``)."""
    weight = torch.load(checkpoint, map_location=map_location, weights_only=False) if isinstance(checkpoint, str) else checkpoint
    if "state_dict" not in weight:
        raise KeyError("checkpoint has no 'state_dict' entry (train.py:288-298 format expected)")
    model.load_state_dict(weight["state_dict"])
    ema = weight.get("ema_state_dict")
    if use_ema and ema is not None:
        copy_parameters(ema["shadow_params"], model.parameters())
    return {"iter": weight.get("iter")}
