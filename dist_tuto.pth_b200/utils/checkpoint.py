"""Checkpoint / resume (absent from the reference).

Rank 0 writes ``{"model": state_dict, "optim": ..., "steps": ..}`` atomically
(tmp file + rename).  Works for ``nn.Module`` models and for the fused trainer
(which exposes ``state_dict()`` over its flat fp32 parameter/momentum buffers
using the reference's parameter names, so checkpoints interchange)."""
from __future__ import annotations

import os
from typing import Any, Dict

import torch

__all__ = ["save_checkpoint", "load_checkpoint"]


def save_checkpoint(path: str, model, optimizer=None, **extra: Any) -> str:
    """Atomically write ``model`` (module or fused trainer), ``optimizer`` state and ``extra`` keys to ``path``."""
    sd = model.state_dict()
    if "model" in sd and isinstance(sd.get("model"), dict):   # fused trainer: already structured
        blob: Dict[str, Any] = dict(sd)
    else:
        blob = {"model": {k: v.detach().cpu() for k, v in sd.items()}}
    if optimizer is not None:
        blob["optim"] = optimizer.state_dict()
    blob.update(extra)
    tmp = f"{path}.tmp.{os.getpid()}"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(blob, tmp)
    os.replace(tmp, path)
    return path


def load_checkpoint(path: str, model=None, optimizer=None, map_location="cpu") -> Dict[str, Any]:
    """Restore ``model`` (an ``nn.Module`` or a fused trainer) and, when given, ``optimizer`` (``torch.optim`` or
    ``FlatSGD``) from ``path``; returns the whole blob (``steps``, ``history``, ...)."""
    blob = torch.load(path, map_location=map_location)
    if model is not None:
        if isinstance(model, torch.nn.Module):
            model.load_state_dict(blob["model"])
        else:                                   # fused trainer: parameters + momentum + step counter in one dict
            model.load_state_dict(blob)
    if optimizer is not None and "optim" in blob:
        optimizer.load_state_dict(blob["optim"])
    return blob
