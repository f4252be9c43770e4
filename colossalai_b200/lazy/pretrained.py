"""Lazy `from_pretrained`: build the model skeleton from a Hugging Face checkpoint directory WITHOUT reading the
weights (parameters stay lazy / meta), remember the path, and let `Booster.boost()` stream each rank's slices in
after the plugin has sharded the model — a 70B checkpoint never has to fit on one device or in one host process.

Parity: reference `colossalai/lazy/pretrained.py:11-328` (`new_from_pretrained` patched over
`PreTrainedModel.from_pretrained` inside `LazyInitContext`, path recorded via `interface/pretrained.py`); we own the
model definitions, so this is a plain function plus a classmethod mixin instead of a monkey patch.
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch

from ..interface.pretrained import set_pretrained_path
from .lazy_init import LazyInitContext

__all__ = ["from_pretrained", "is_hf_checkpoint_dir", "load_pretrained_into"]


def is_hf_checkpoint_dir(path: str) -> bool:
    """A directory with an HF `config.json` (it names a `model_type`) next to `*.safetensors` / `*.bin` shards."""
    cfg = os.path.join(path, "config.json")
    if not (os.path.isdir(path) and os.path.isfile(cfg)):
        return False
    try:
        with open(cfg) as f:
            return "model_type" in json.load(f)
    except (OSError, ValueError):
        return False


def from_pretrained(path: str, lazy: bool = True, dtype: Optional[torch.dtype] = None, **config_overrides):
    """`lazy=True`: skeleton only (weights load inside `Booster.boost`); `lazy=False`: load right away."""
    from ..models import build_model
    from ..models.hf_io import config_from_hf, load_hf_checkpoint

    if not lazy:
        return load_hf_checkpoint(path, dtype=dtype or torch.bfloat16)
    with open(os.path.join(path, "config.json")) as f:
        cfg = config_from_hf(json.load(f))
    if config_overrides:
        cfg = cfg.replace(**config_overrides)
    with LazyInitContext():
        model = build_model(cfg)
    if dtype is not None:
        model = model.to(dtype)
    set_pretrained_path(model, path)
    return model


def load_pretrained_into(model: torch.nn.Module, path: str, strict: bool = False) -> None:
    """Fill an already sharded / wrapped model from an HF directory: every rank converts names (fused qkv, gate|up,
    stacked experts) and keeps its own tensor-/expert-parallel slice of each weight."""
    from ..inference.core.plugin import InferCheckpoint_io
    from ..interface import ModelWrapper

    inner = model.unwrap() if isinstance(model, ModelWrapper) else model
    InferCheckpoint_io(verbose=False).load_model(inner, path, strict=strict)
