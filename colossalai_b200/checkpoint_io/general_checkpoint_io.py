"""Single-process / pure-DP checkpoint IO (HF-compatible sharded layout).
Parity: reference `colossalai/checkpoint_io/general_checkpoint_io.py:37-300`."""
from __future__ import annotations

import gc
import logging
import os
from functools import reduce
from pathlib import Path
from typing import Optional

import torch
import torch.nn as nn
from torch.optim import Optimizer

from ..interface import ModelWrapper, OptimizerWrapper
from .checkpoint_io_base import CheckpointIO
from .index_file import CheckpointIndexFile
from .utils import (
    async_save_state_dict_shards,
    get_model_base_filenames,
    get_optimizer_base_filenames,
    is_safetensor_checkpoint,
    load_param_groups_into_optimizer,
    load_shard_state_dict,
    load_state_dict,
    load_state_dict_into_model,
    load_states_into_optimizer,
    save_param_groups,
    save_state_dict,
    save_state_dict_shards,
    shard_model_checkpoint,
    shard_optimizer_checkpoint,
    sharded_optimizer_loading_epilogue,
    unflatten_optim_state,
)

__all__ = ["GeneralCheckpointIO"]


def _unwrap(model):
    return model.unwrap() if isinstance(model, ModelWrapper) else model


def _unwrap_optim(optimizer):
    return optimizer.unwrap() if isinstance(optimizer, OptimizerWrapper) else optimizer


class GeneralCheckpointIO(CheckpointIO):
    # ------------------------------------------------------------------ model
    def load_unsharded_model(self, model: nn.Module, checkpoint: str, strict: bool, low_cpu_mem_mode: bool = True,
                             num_threads: int = 1):
        _unwrap(model).load_state_dict(load_state_dict(checkpoint), strict=strict)

    def save_unsharded_model(self, model: nn.Module, checkpoint: str, gather_dtensor: bool, use_safetensors: bool,
                             use_async: bool = False):
        sd = _unwrap(model).state_dict()
        if use_async:
            from ..utils.aio import AsyncSafetensorsWriter

            cache = self.pinned_state_dicts.setdefault(hash(model), {})
            staged = {}
            for k, t in sd.items():
                buf = cache.get(k)
                if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                    buf = cache[k] = torch.empty(t.shape, dtype=t.dtype, device="cpu",
                                                 pin_memory=torch.cuda.is_available())
                buf.copy_(t.detach(), non_blocking=True)
                staged[k] = buf
            self._sync_d2h()
            w = AsyncSafetensorsWriter(checkpoint)
            w.write(staged)
            self.async_writers.append(w)
        else:
            save_state_dict(sd, checkpoint, use_safetensors)

    def save_sharded_model(self, model: nn.Module, checkpoint_path: str, gather_dtensor: bool = False,
                           prefix: Optional[str] = None, max_shard_size: int = 1024, use_safetensors: bool = False,
                           use_async: bool = False):
        if os.path.isfile(checkpoint_path):
            logging.error(f"Provided path ({checkpoint_path}) should be a directory, not a file")
            return
        Path(checkpoint_path).mkdir(parents=True, exist_ok=True)
        sd = _unwrap(model).state_dict()
        shards = shard_model_checkpoint(sd, max_shard_size=max_shard_size)
        weights_name, save_index_file = get_model_base_filenames(prefix, use_safetensors)
        index_file = CheckpointIndexFile(checkpoint_path)
        if use_async:
            cache = self.pinned_state_dicts.setdefault(hash(model), {})
            total = async_save_state_dict_shards(shards, checkpoint_path, index_file, weights_name, True, cache,
                                                 self.async_writers)
        else:
            total = save_state_dict_shards(shards, checkpoint_path, index_file, weights_name, True, use_safetensors)
        index_file.append_meta_data("total_size", total)
        index_file.write_index_file(save_index_file)
        cfg = getattr(_unwrap(model), "config", None)
        if cfg is not None and hasattr(cfg, "to_dict"):
            import json

            with open(os.path.join(checkpoint_path, "config.json"), "w") as f:
                json.dump(cfg.to_dict(), f, indent=2, default=str)

    def load_sharded_model(self, model: nn.Module, checkpoint_index_file: Path, strict: bool = False,
                           use_safetensors: bool = False, load_sub_module: bool = True,
                           low_cpu_mem_mode: bool = True, num_threads: int = 1):
        model = _unwrap(model)
        ckpt_index = CheckpointIndexFile.from_file(checkpoint_index_file)
        missing_lists = []
        for shard_file in ckpt_index.get_checkpoint_filenames():
            sd = load_shard_state_dict(Path(shard_file), is_safetensor_checkpoint(shard_file))
            load_state_dict_into_model(model, sd, missing_lists, strict, load_sub_module)
            del sd
            gc.collect()
        if strict and missing_lists:
            remain = reduce(lambda a, b: set(a) & set(b), missing_lists)
            if remain:
                raise RuntimeError("Error(s) in loading state_dict for {}:\n\tMissing key(s): {}".format(
                    model.__class__.__name__, ", ".join(f'"{k}"' for k in sorted(remain))))

    # ------------------------------------------------------------------ optimizer
    def save_unsharded_optimizer(self, optimizer: Optimizer, checkpoint: Path, gather_dtensor: bool,
                                 use_async: bool = False):
        torch.save(_unwrap_optim(optimizer).state_dict(), checkpoint)

    def load_unsharded_optimizer(self, optimizer: Optimizer, checkpoint: Path, low_cpu_mem_mode: bool = True,
                                 num_threads: int = 1):
        _unwrap_optim(optimizer).load_state_dict(load_state_dict(checkpoint))

    def save_sharded_optimizer(self, optimizer: Optimizer, checkpoint: Path, gather_dtensor: bool, prefix: str,
                               size_per_shard: int, use_async: bool = False):
        optimizer = _unwrap_optim(optimizer)
        if os.path.isfile(checkpoint):
            logging.error(f"Provided path ({checkpoint}) should be a directory, not a file")
            return
        Path(checkpoint).mkdir(parents=True, exist_ok=True)
        sd = optimizer.state_dict()
        shards = shard_optimizer_checkpoint(sd, max_shard_size=size_per_shard)
        states_name, save_index_file, param_group_file = get_optimizer_base_filenames(prefix)
        index_file = CheckpointIndexFile(checkpoint)
        index_file.append_meta_data("param_groups", param_group_file)
        save_param_groups(sd, os.path.join(checkpoint, param_group_file))
        total = save_state_dict_shards(shards, checkpoint, index_file, states_name, True, use_safetensors=False)
        index_file.append_meta_data("total_size", total)
        index_file.write_index_file(save_index_file)

    def load_sharded_optimizer(self, optimizer: Optimizer, index_file_path: str, prefix: str,
                               low_cpu_mem_mode: bool = True, num_threads: int = 1):
        optimizer = _unwrap_optim(optimizer)
        ckpt_index = CheckpointIndexFile.from_file(index_file_path)
        pg_path = ckpt_index.get_param_group_filename()
        if pg_path is None:
            raise RuntimeError(f"Invalid index file path {index_file_path}: it has no param_groups entry")
        id_map = load_param_groups_into_optimizer(optimizer, pg_path)
        for shard_file in ckpt_index.get_checkpoint_filenames():
            sd = load_shard_state_dict(Path(shard_file))
            if is_safetensor_checkpoint(shard_file):
                sd = unflatten_optim_state(sd)
            load_states_into_optimizer(optimizer, sd, id_map)
        sharded_optimizer_loading_epilogue(optimizer)
