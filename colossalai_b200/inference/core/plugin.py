"""Checkpoint IO of the inference engine: HF checkpoints straight into TP-sharded inference modules.
Parity: reference `colossalai/inference/core/plugin.py:21-140` (`InferCheckpoint_io`)."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from ...checkpoint_io.hybrid_parallel_checkpoint_io import _padded_like
from ...models.hf_io import config_from_hf, convert_hf_state_dict, iter_hf_shards
from ...tensor.d_tensor import is_distributed_tensor
from ...tensor.d_tensor.api import distribute_tensor_with_spec

__all__ = ["InferCheckpoint_io"]


class InferCheckpoint_io:
    """`load_model(model, checkpoint_dir)`: every rank reads the HF shards, converts names / fuses q-k-v and gate-up,
    and keeps only its own tensor-parallel slice of each weight."""

    def __init__(self, verbose: bool = True) -> None:
        self.verbose = verbose

    @staticmethod
    def _assign(param: torch.Tensor, full: torch.Tensor) -> None:
        t = _padded_like(param, full)
        if is_distributed_tensor(param) and tuple(t.shape) != tuple(param.shape):
            t = distribute_tensor_with_spec(t, param)
        assert tuple(t.shape) == tuple(param.shape), f"{tuple(t.shape)} vs {tuple(param.shape)}"
        with torch.no_grad():
            param.copy_(t.to(device=param.device, dtype=param.dtype))

    def load_model(self, model: nn.Module, checkpoint: str, strict: bool = False) -> nn.Module:
        with open(os.path.join(checkpoint, "config.json")) as f:
            cfg = config_from_hf(json.load(f))
        merged: Dict[str, torch.Tensor] = {}
        for shard in iter_hf_shards(checkpoint):
            merged.update(shard)
        sd = convert_hf_state_dict(merged, cfg)
        params = dict(model.named_parameters())
        missing = []
        for name, p in params.items():
            if name in sd:
                self._assign(p, sd[name])
            elif not (cfg.tie_word_embeddings and name == "lm_head.weight"):
                missing.append(name)
        if strict and missing:
            raise RuntimeError(f"InferCheckpoint_io: missing keys {missing[:8]}")
        return model

    def save_model(self, model: nn.Module, checkpoint: str, use_safetensors: bool = True) -> None:
        """Gather TP shards and write one HF-style file (rank 0 only)."""
        import torch.distributed as dist

        from ...models.hf_io import to_hf_state_dict
        from ...tensor.d_tensor import to_global

        full = {n: to_global(p).detach().cpu() for n, p in model.named_parameters()}
        if dist.is_initialized() and dist.get_rank() != 0:
            return
        os.makedirs(checkpoint, exist_ok=True)

        class _Holder:
            cfg = model.cfg

            @staticmethod
            def state_dict():
                return full

        sd = to_hf_state_dict(_Holder, model.cfg)
        if use_safetensors:
            from safetensors.torch import save_file

            save_file({k: v.contiguous() for k, v in sd.items()}, os.path.join(checkpoint, "model.safetensors"))
        else:
            torch.save(sd, os.path.join(checkpoint, "pytorch_model.bin"))
