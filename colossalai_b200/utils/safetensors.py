"""Asynchronous safetensors checkpoints through the native AIO writer.

Parity: reference `colossalai/utils/safetensors.py` (`save`, `save_nested`, `move_and_save`, `load_flat`,
`create_pinned_state_dict`, `_flatten_optim_state_dict` / `_unflatten_optim_state_dict`).  A device state dict is
copied into (reusable) pinned host buffers, then handed to `AsyncSafetensorsWriter` (`kernel/csrc/async_file_io.cpp`,
a pthread pool) so the training loop only pays the D2H copy; the returned writer is joined with `.synchronize()`.
Nested optimizer states are flattened to `state.<param id>.<name>` keys, non-tensor leaves and the param groups travel
as JSON in the safetensors metadata.
"""
from __future__ import annotations

import json
from typing import Any, Dict, Optional, Tuple

import torch

from .aio import AsyncSafetensorsWriter

__all__ = ["save", "save_nested", "move_and_save", "load_flat", "create_pinned_state_dict", "flatten_optim_state_dict",
           "unflatten_optim_state_dict"]

_SEP = "."
_TENSOR_FLAG = "__tensor__"


def create_pinned_state_dict(state_dict: Dict[str, Any], empty: bool = True, num_threads: int = 1) -> Dict[str, Any]:
    """Host mirror of `state_dict` (nested dicts allowed) in page-locked memory when a CUDA device is present."""
    pin = torch.cuda.is_available()

    def mirror(v):
        if isinstance(v, dict):
            return {k: mirror(x) for k, x in v.items()}
        if torch.is_tensor(v):
            buf = torch.empty(v.shape, dtype=v.dtype, device="cpu", pin_memory=pin)
            if not empty:
                buf.copy_(v)
            return buf
        return v

    return mirror(state_dict)


def flatten_optim_state_dict(optim_state_dict: Dict[str, Any], separator: str = _SEP
                             ) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    """-> (flat tensors, metadata).  Tensors keep their place under `state.<id>.<name>`; everything else (ints,
    floats, bools, param_groups) is serialised into the metadata."""
    flat: Dict[str, torch.Tensor] = {}
    other: Dict[str, Any] = {}

    def walk(prefix: str, v) -> None:
        if isinstance(v, dict):
            for k, x in v.items():
                walk(f"{prefix}{separator}{k}" if prefix else str(k), x)
        elif torch.is_tensor(v):
            flat[prefix] = v
        else:
            other[prefix] = v

    walk("state", optim_state_dict.get("state", {}))
    meta = {"non_tensor": json.dumps(other), "param_groups": json.dumps(optim_state_dict.get("param_groups", [])),
            "format": "pt"}
    return flat, meta


def unflatten_optim_state_dict(flat: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None,
                               separator: str = _SEP) -> Dict[str, Any]:
    out: Dict[str, Any] = {"state": {}, "param_groups": []}
    leaves: Dict[str, Any] = dict(flat)
    if metadata:
        leaves.update(json.loads(metadata.get("non_tensor", "{}")))
        out["param_groups"] = json.loads(metadata.get("param_groups", "[]"))
    for key, v in leaves.items():
        parts = key.split(separator)
        assert parts[0] == "state", key
        pid: Any = int(parts[1]) if parts[1].lstrip("-").isdigit() else parts[1]
        out["state"].setdefault(pid, {})[separator.join(parts[2:])] = v
    return out


def _stage(state_dict: Dict[str, torch.Tensor], pinned: Optional[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    staged = {}
    for k, v in state_dict.items():
        if pinned is not None and k in pinned:
            pinned[k].copy_(v, non_blocking=True)
            staged[k] = pinned[k]
        else:
            staged[k] = v.detach().to("cpu", non_blocking=True) if v.device.type != "cpu" else v.detach()
    if torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()
    return staged


def save(path: str, state_dict: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None,
         n_threads: int = 4) -> AsyncSafetensorsWriter:
    """Start writing host tensors to `path`; returns the writer (call `.synchronize()` before reading the file)."""
    w = AsyncSafetensorsWriter(path, n_threads=n_threads)
    w.write({k: v for k, v in state_dict.items()}, metadata)
    return w


def move_and_save(path: str, state_dict: Dict[str, torch.Tensor],
                  state_dict_pinned: Optional[Dict[str, torch.Tensor]] = None,
                  metadata: Optional[Dict[str, str]] = None) -> AsyncSafetensorsWriter:
    """Device tensors -> pinned host buffers -> background write."""
    return save(path, _stage(state_dict, state_dict_pinned), metadata)


def save_nested(path: str, optim_state_dict: Dict[str, Any],
                state_dict_pinned: Optional[Dict[str, torch.Tensor]] = None) -> AsyncSafetensorsWriter:
    flat, meta = flatten_optim_state_dict(optim_state_dict)
    pinned_flat = None
    if state_dict_pinned is not None:
        pinned_flat, _ = flatten_optim_state_dict(state_dict_pinned) if "state" in state_dict_pinned else \
            (state_dict_pinned, None)
    return save(path, _stage(flat, pinned_flat), meta)


def load_flat(checkpoint_path: str, seperator: str = _SEP) -> Dict[str, Any]:
    """Read a file written by `save_nested` back into `{state: {id: {...}}, param_groups: [...]}`."""
    from safetensors import safe_open

    flat: Dict[str, torch.Tensor] = {}
    with safe_open(checkpoint_path, framework="pt") as f:
        meta = f.metadata()
        for k in f.keys():
            flat[k] = f.get_tensor(k)
    return unflatten_optim_state_dict(flat, meta, seperator)
