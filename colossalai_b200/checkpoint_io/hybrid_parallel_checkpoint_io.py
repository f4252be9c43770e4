"""Checkpoint IO for TP x PP x SP x DP(ZeRO) models.

Parity: reference `colossalai/checkpoint_io/hybrid_parallel_checkpoint_io.py:59-1167`: save gathers TP shards and
strips vocab padding so the files are HF-layout and parallelism-agnostic; only (dp 0, sp 0, tp 0) ranks write; each
PP stage writes its own shard files + a stage-local index which the master merges; load re-shards every tensor with
the parameter's own sharding function; optimizer states are gathered over TP (and over DP for ZeRO) likewise.
"""
from __future__ import annotations

import copy
import logging
import os
from collections import OrderedDict
from pathlib import Path
from shutil import rmtree
from typing import Dict, Iterator, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.distributed import ProcessGroup
from torch.optim import Optimizer

from ..interface import ModelWrapper, OptimizerWrapper
from ..parallel import comm
from ..tensor.d_tensor import distribute_tensor_with_spec, is_distributed_tensor, to_global
from ..tensor.padded_tensor import is_padded_tensor, to_padded_tensor, to_unpadded_tensor
from .general_checkpoint_io import GeneralCheckpointIO
from .index_file import CheckpointIndexFile
from .utils import (
    StateDictSharder,
    async_save_state_dict_shards,
    gather_distributed_param,
    get_model_base_filenames,
    get_optimizer_base_filenames,
    is_safetensor_checkpoint,
    load_shard_state_dict,
    load_state_dict,
    save_param_groups,
    save_state_dict,
    save_state_dict_shards,
)

__all__ = ["HybridParallelCheckpointIO"]


def _padded_like(param: torch.Tensor, tensor: torch.Tensor) -> torch.Tensor:
    """Pad a GLOBAL un-padded tensor so it can be sharded like `param` (vocab padding)."""
    gshape = getattr(param, "dist_global_shape", None)
    if gshape is not None and tuple(tensor.shape) != tuple(gshape):
        for d, (a, b) in enumerate(zip(tensor.shape, gshape)):
            if a < b:
                pad_shape = list(tensor.shape)
                pad_shape[d] = b - a
                tensor = torch.cat([tensor, tensor.new_zeros(pad_shape)], dim=d)
    return tensor


def _unpadded_global(param: torch.Tensor, module_of: Dict[int, nn.Module]) -> torch.Tensor:
    """TP-gather a parameter and strip vocab padding using the owning module's `old_num_embeddings`."""
    full = to_global(param) if is_distributed_tensor(param) else param.detach()
    mod = module_of.get(id(param))
    old = getattr(mod, "old_num_embeddings", None) if mod is not None else None
    if old is not None and full.dim() >= 1 and full.shape[0] > old and param is getattr(mod, "weight", None):
        full = full[:old]
    if old is not None and full.dim() == 1 and full.shape[0] > old and param is getattr(mod, "bias", None):
        full = full[:old]
    return full


class HybridParallelCheckpointIO(GeneralCheckpointIO):
    def __init__(self, dp_group: ProcessGroup, pp_group: ProcessGroup, tp_group: ProcessGroup,
                 sp_group: ProcessGroup, zero_stage: int, verbose: bool = True) -> None:
        super().__init__()
        self.global_dp_group, self.pp_group, self.tp_group, self.sp_group = dp_group, pp_group, tp_group, sp_group
        self.dp_rank = dist.get_rank(dp_group) if dp_group is not None else 0
        self.tp_rank = dist.get_rank(tp_group) if tp_group is not None else 0
        self.pp_rank = dist.get_rank(pp_group) if pp_group is not None else 0
        self.sp_rank = dist.get_rank(sp_group) if sp_group is not None else 0
        self.dp_size = dist.get_world_size(dp_group) if dp_group is not None else 1
        self.pp_size = dist.get_world_size(pp_group) if pp_group is not None else 1
        self.tp_size = dist.get_world_size(tp_group) if tp_group is not None else 1
        self.use_zero = zero_stage > 0
        self.verbose = verbose
        self.coordinator_is_master = dist.get_rank() == 0

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _module_of(model: nn.Module) -> Dict[int, nn.Module]:
        return {id(p): m for m in model.modules() for p in m._parameters.values() if p is not None}

    def _is_writer(self) -> bool:
        sp_writer = self.sp_rank == 0 or self.sp_group is self.tp_group
        return self.dp_rank == 0 and self.tp_rank == 0 and sp_writer

    def _model_sharder(self, model: nn.Module, size_per_shard: int = 1024) -> Iterator[Tuple[OrderedDict, float]]:
        """Yield (state-dict shard, size MB) with TP-gathered, un-padded tensors.  Collective over the TP group."""
        sharder = StateDictSharder(size_per_shard)
        module_of = self._module_of(model)
        seen = set()
        for name, param in model.named_parameters():
            if param is None or id(param) in seen:
                continue
            seen.add(id(param))
            full = _unpadded_global(param, module_of)
            block, size = sharder.append_param(name, full)
            if block is not None:
                yield block, size
        # tied parameters appear under several names in a state dict
        name_of = {}
        for name, param in model.named_parameters(remove_duplicate=False):
            if id(param) in name_of and name != name_of[id(param)]:
                pass
            name_of.setdefault(id(param), name)
        for name, buf in model.named_buffers():
            if buf is None:
                continue
            mod_path, _, bname = name.rpartition(".")
            owner = model.get_submodule(mod_path) if mod_path else model
            if bname in owner._non_persistent_buffers_set:
                continue
            block, size = sharder.append_param(name, buf.detach())
            if block is not None:
                yield block, size
        yield sharder.current_block, sharder.current_block_size

    # ------------------------------------------------------------------ model
    def save_sharded_model(self, model: ModelWrapper, checkpoint: str, gather_dtensor: bool = True,
                           prefix: Optional[str] = None, size_per_shard: int = 1024, use_safetensors: bool = False,
                           use_async: bool = False) -> None:
        model = model.unwrap() if isinstance(model, ModelWrapper) else model
        if os.path.isfile(checkpoint):
            logging.error(f"Provided path ({checkpoint}) should be a directory, not a file")
            return
        Path(checkpoint).mkdir(parents=True, exist_ok=True)
        weights_name, save_index_file = get_model_base_filenames(prefix, use_safetensors or use_async)
        index_file = CheckpointIndexFile(checkpoint)
        shards = self._model_sharder(model, size_per_shard)
        writer = self._is_writer()
        if self.pp_size == 1:
            if use_async:
                total = async_save_state_dict_shards(shards, checkpoint, index_file, weights_name, writer,
                                                     self.pinned_state_dicts.setdefault(id(model), {}),
                                                     self.async_writers)
            else:
                total = save_state_dict_shards(shards, checkpoint, index_file, weights_name, writer,
                                               use_safetensors)
            if writer:
                index_file.append_meta_data("total_size", total)
                index_file.write_index_file(save_index_file)
                self._save_config(model, checkpoint)
        else:
            # every stage writes files tagged with its stage id + a temporary stage index; master merges
            tmp_dir = os.path.join(checkpoint, "tmp_index_files")
            Path(tmp_dir).mkdir(parents=True, exist_ok=True)
            stage_weights = weights_name.replace(".bin", f"-stage-{self.pp_rank + 1:05d}-shard.bin") \
                .replace(".safetensors", f"-stage-{self.pp_rank + 1:05d}-shard.safetensors")
            stage_index = save_index_file.replace(".json", f"-stage-{self.pp_rank + 1:05d}.json")
            if use_async:
                total = async_save_state_dict_shards(shards, checkpoint, index_file, stage_weights, writer,
                                                     self.pinned_state_dicts.setdefault(id(model), {}),
                                                     self.async_writers)
            else:
                total = save_state_dict_shards(shards, checkpoint, index_file, stage_weights, writer,
                                               use_safetensors)
            if writer:
                index_file.append_meta_data("total_size", total)
                index_file.export(os.path.join(tmp_dir, stage_index))
            dist.barrier(self.pp_group)
            if writer and self.pp_rank == 0:
                final = CheckpointIndexFile(checkpoint)
                final.append_meta_data("total_size", 0.0)
                for fn in sorted(os.listdir(tmp_dir)):
                    st = CheckpointIndexFile.from_file(os.path.join(tmp_dir, fn))
                    final.metadata["total_size"] += st.metadata.get("total_size", 0.0)
                    for k, v in st.weight_map.items():
                        final.append_weight_map(k, v)
                final.write_index_file(save_index_file)
                self._save_config(model, checkpoint)
                rmtree(tmp_dir, ignore_errors=True)
        if dist.is_initialized():
            dist.barrier()

    @staticmethod
    def _save_config(model: nn.Module, checkpoint: str) -> None:
        cfg = getattr(model, "config", None)
        if cfg is not None and hasattr(cfg, "to_dict"):
            import json

            with open(os.path.join(checkpoint, "config.json"), "w") as f:
                json.dump(cfg.to_dict(), f, indent=2, default=str)

    def _load_param_from(self, param: torch.Tensor, tensor: torch.Tensor) -> None:
        tensor = _padded_like(param, tensor)
        if is_distributed_tensor(param) and tuple(tensor.shape) != tuple(param.shape):
            tensor = distribute_tensor_with_spec(tensor, param)
        assert tuple(tensor.shape) == tuple(param.shape), (
            f"checkpoint tensor {tuple(tensor.shape)} cannot be mapped onto parameter {tuple(param.shape)}")
        with torch.no_grad():
            param.copy_(tensor.to(param.dtype))

    def load_sharded_model(self, model: ModelWrapper, checkpoint_index_file: Path, strict: bool = False,
                           low_cpu_mem_mode: bool = True, num_threads: int = 1) -> None:
        model_before = model
        model = model.unwrap() if isinstance(model, ModelWrapper) else model
        ckpt_index = CheckpointIndexFile.from_file(checkpoint_index_file)
        weight_map = ckpt_index.weight_map
        root = Path(ckpt_index.root_path)
        by_file: Dict[str, List[str]] = {}
        params = dict(model.named_parameters(remove_duplicate=False))
        buffers = dict(model.named_buffers())
        missing = []
        for name in list(params.keys()) + list(buffers.keys()):
            if name in weight_map:
                by_file.setdefault(weight_map[name], []).append(name)
            elif name in params:
                missing.append(name)
        loaded = set()
        for fn, names in by_file.items():
            sd = load_shard_state_dict(root / fn, is_safetensor_checkpoint(fn))
            for name in names:
                t = sd[name]
                if name in params:
                    if id(params[name]) in loaded:
                        continue
                    self._load_param_from(params[name], t)
                    loaded.add(id(params[name]))
                else:
                    with torch.no_grad():
                        buffers[name].copy_(t)
            del sd
        missing = [m for m in missing if id(params[m]) not in loaded]
        if strict and missing:
            raise RuntimeError(f"Error(s) in loading state_dict for {model.__class__.__name__}:\n\tMissing key(s): "
                               + ", ".join(missing))
        if isinstance(model_before, ModelWrapper) and hasattr(model_before, "update_master_params"):
            model_before.update_master_params()

    def save_unsharded_model(self, model: ModelWrapper, checkpoint: str, gather_dtensor: bool, use_safetensors: bool,
                             use_async: bool = False) -> None:
        model = model.unwrap() if isinstance(model, ModelWrapper) else model
        sd = OrderedDict()
        for block, _ in self._model_sharder(model, size_per_shard=1 << 40):
            sd.update(block)
        if self.pp_size > 1:
            gathered = [None] * self.pp_size
            dist.all_gather_object(gathered, {k: v.cpu() for k, v in sd.items()}, group=self.pp_group)
            sd = OrderedDict()
            for part in gathered:
                sd.update(part)
        if self._is_writer() and self.pp_rank == 0:
            save_state_dict(sd, checkpoint, use_safetensors)
        if dist.is_initialized():
            dist.barrier()

    def load_unsharded_model(self, model: ModelWrapper, checkpoint: str, strict: bool = False,
                             low_cpu_mem_mode: bool = True, num_threads: int = 1) -> None:
        model_before = model
        model = model.unwrap() if isinstance(model, ModelWrapper) else model
        sd = load_state_dict(checkpoint)
        params = dict(model.named_parameters(remove_duplicate=False))
        missing = []
        for name, p in params.items():
            if name in sd:
                self._load_param_from(p, sd[name])
            else:
                missing.append(name)
        for name, b in model.named_buffers():
            if name in sd:
                with torch.no_grad():
                    b.copy_(sd[name])
        if strict and missing:
            raise RuntimeError(f"Missing key(s) in state_dict: {missing}")
        if isinstance(model_before, ModelWrapper) and hasattr(model_before, "update_master_params"):
            model_before.update_master_params()

    # ------------------------------------------------------------------ optimizer
    def _optim_state_global(self, optimizer: OptimizerWrapper, model: nn.Module) -> Tuple[Dict[int, Dict], List[dict]]:
        """{param_id: {state_name: global tensor}} keyed by the ORIGINAL param ids (param_info), TP-gathered."""
        optim = optimizer.unwrap() if isinstance(optimizer, OptimizerWrapper) else optimizer
        m2w = getattr(optimizer, "master_to_working_map", {}) or {}
        param_info = getattr(optimizer, "param_info", None) or {}
        module_of = self._module_of(model)
        name_of = {id(p): n for n, p in model.named_parameters()}
        states: Dict[int, Dict] = {}
        pid = 0
        groups = []
        if hasattr(optimizer, "working_params_in_state_order"):
            # ZeRO: states live in dp-sharded flat buckets.  `state_dict()` gathers them over dp and splits them back
            # per working parameter (still TP / EP local); globalise those exactly like the parameters themselves.
            sd = optimizer.state_dict()
            for i, wp in enumerate(optimizer.working_params_in_state_order()):
                out = {}
                for k, v in sd["state"].get(i, {}).items():
                    if torch.is_tensor(v) and v.dim() > 0 and tuple(v.shape) == tuple(wp.shape):
                        v = v.to(wp.device)
                        for a in ("dist_shard", "shard_fn", "gather_fn", "dist_global_shape"):
                            if hasattr(wp, a):
                                try:
                                    setattr(v, a, getattr(wp, a))
                                except Exception:
                                    pass
                        full = to_global(v) if is_distributed_tensor(wp) else v.detach()
                        mod = module_of.get(id(wp))
                        old = getattr(mod, "old_num_embeddings", None) if mod is not None else None
                        if old is not None and full.shape[0] > old:
                            full = full[:old]
                        out[k] = full.cpu()
                    else:
                        out[k] = v.detach().cpu() if torch.is_tensor(v) else v
                out["__name__"] = name_of.get(id(wp), str(i))
                states[i] = out
            return states, sd["param_groups"]
        for g in optim.param_groups:
            ids = []
            for mp in g["params"]:
                wp = m2w.get(mp, mp)
                st = optimizer.get_full_state(mp) if hasattr(optimizer, "get_full_state") else optim.state.get(mp, {})
                out = {}
                for k, v in st.items():
                    if torch.is_tensor(v) and v.dim() > 0 and v.shape == mp.shape:
                        for a in ("dist_shard", "shard_fn", "gather_fn", "dist_global_shape"):
                            if hasattr(wp, a):
                                try:
                                    setattr(v, a, getattr(wp, a))
                                except Exception:
                                    pass
                        full = to_global(v) if is_distributed_tensor(wp) else v.detach()
                        mod = module_of.get(id(wp))
                        old = getattr(mod, "old_num_embeddings", None) if mod is not None else None
                        if old is not None and full.shape[0] > old:
                            full = full[:old]
                        out[k] = full.cpu()
                    elif torch.is_tensor(v):
                        out[k] = v.detach().cpu()
                    else:
                        out[k] = v
                out["__name__"] = name_of.get(id(wp), str(pid))
                states[pid] = out
                ids.append(pid)
                pid += 1
            groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": ids})
        return states, groups

    def save_sharded_optimizer(self, optimizer: OptimizerWrapper, checkpoint: str, gather_dtensor: bool = True,
                               prefix: Optional[str] = None, size_per_shard: int = 1024, use_async: bool = False):
        assert isinstance(optimizer, OptimizerWrapper), "Please boost the optimizer before saving!"
        model = optimizer.model.unwrap() if hasattr(optimizer, "model") else None
        Path(checkpoint).mkdir(parents=True, exist_ok=True)
        states, groups = self._optim_state_global(optimizer, model)
        if self.pp_size > 1:
            gathered = [None] * self.pp_size
            dist.all_gather_object(gathered, (states, groups), group=self.pp_group)
            # params are keyed by name across stages
            states = {}
            for st, _ in gathered:
                for _, v in st.items():
                    states[len(states)] = v
        if self._is_writer() and self.pp_rank == 0:
            states_name, save_index_file, param_group_file = get_optimizer_base_filenames(prefix)
            index_file = CheckpointIndexFile(checkpoint)
            index_file.append_meta_data("param_groups", param_group_file)
            torch.save(groups, os.path.join(checkpoint, param_group_file))
            sharder = StateDictSharder(size_per_shard)

            def gen():
                for pid, st in states.items():
                    block, size = sharder.append_optim_state(pid, st)
                    if block is not None:
                        yield block, size
                yield sharder.current_block, sharder.current_block_size

            total = save_state_dict_shards(gen(), checkpoint, index_file, states_name, True, use_safetensors=False)
            index_file.append_meta_data("total_size", total)
            index_file.write_index_file(save_index_file)
        if dist.is_initialized():
            dist.barrier()

    def load_sharded_optimizer(self, optimizer: OptimizerWrapper, checkpoint_index_file: str, prefix: str = "",
                               low_cpu_mem_mode: bool = True, num_threads: int = 1):
        assert isinstance(optimizer, OptimizerWrapper), "Please boost the optimizer before loading!"
        model = optimizer.model.unwrap()
        optim = optimizer.unwrap()
        ckpt_index = CheckpointIndexFile.from_file(checkpoint_index_file)
        all_states: Dict[str, Dict] = {}
        for fn in ckpt_index.get_checkpoint_filenames():
            sd = load_shard_state_dict(Path(fn))
            for _, st in sd.items():
                all_states[st.get("__name__")] = st
        self._load_optim_states(optimizer, model, all_states)
        pg_file = ckpt_index.get_param_group_filename()
        if pg_file is not None and os.path.exists(pg_file):
            self._load_param_group_hparams(optim, torch.load(pg_file, weights_only=False))

    @staticmethod
    def _load_param_group_hparams(optim, saved_groups) -> None:
        """Restore the per-group hyper-parameters and counters (lr, betas, weight decay, the fused optimizers' shared
        `step`) — everything of a param group except the parameter list, which belongs to the current layout.  The
        number of groups is layout independent (PP stages and TP ranks all build the same groups)."""
        if isinstance(saved_groups, dict):
            saved_groups = saved_groups.get("param_groups", saved_groups)
        for g, sg in zip(optim.param_groups, saved_groups):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v

    def _load_optim_states(self, optimizer, model, by_name: Dict[str, Dict]) -> None:
        if hasattr(optimizer, "working_params_in_state_order"):          # ZeRO flat buckets
            name_of = {id(p): n for n, p in model.named_parameters()}
            local: Dict[int, Dict] = {}
            for i, wp in enumerate(optimizer.working_params_in_state_order()):
                st = by_name.get(name_of.get(id(wp)))
                if st is None:
                    continue
                entry = {}
                for k, v in st.items():
                    if k == "__name__":
                        continue
                    if torch.is_tensor(v) and v.dim() > 0:
                        v = _padded_like(wp, v)
                        if is_distributed_tensor(wp) and tuple(v.shape) != tuple(wp.shape):
                            v = distribute_tensor_with_spec(v, wp)
                    entry[k] = v
                local[i] = entry
            groups = [{k: v for k, v in g.items() if k != "params"} for g in optimizer.unwrap().param_groups]
            optimizer.load_state_dict({"state": local, "param_groups": groups})
            return
        optim = optimizer.unwrap()
        m2w = getattr(optimizer, "master_to_working_map", {}) or {}
        name_of = {id(p): n for n, p in model.named_parameters()}
        for g in optim.param_groups:
            for mp in g["params"]:
                wp = m2w.get(mp, mp)
                st = by_name.get(name_of.get(id(wp)))
                if st is None:
                    continue
                new = {}
                for k, v in st.items():
                    if k == "__name__":
                        continue
                    if torch.is_tensor(v) and v.dim() > 0:
                        v = _padded_like(wp, v)
                        if is_distributed_tensor(wp) and tuple(v.shape) != tuple(mp.shape):
                            v = distribute_tensor_with_spec(v, wp)
                        v = v.to(mp.device)
                    new[k] = v
                if hasattr(optimizer, "set_full_state"):
                    optimizer.set_full_state(mp, new)
                else:
                    optim.state[mp] = new

    def save_unsharded_optimizer(self, optimizer: OptimizerWrapper, checkpoint: str, gather_dtensor: bool,
                                 use_async: bool = False):
        model = optimizer.model.unwrap()
        states, groups = self._optim_state_global(optimizer, model)
        if self.pp_size > 1:
            gathered = [None] * self.pp_size
            dist.all_gather_object(gathered, states, group=self.pp_group)
            states = {}
            for st in gathered:
                for _, v in st.items():
                    states[len(states)] = v
        if self._is_writer() and self.pp_rank == 0:
            torch.save({"state": states, "param_groups": groups}, checkpoint)
        if dist.is_initialized():
            dist.barrier()

    def load_unsharded_optimizer(self, optimizer: OptimizerWrapper, checkpoint: str, low_cpu_mem_mode: bool = True,
                                 num_threads: int = 1):
        sd = load_state_dict(checkpoint)
        by_name = {st.get("__name__"): st for st in sd["state"].values()}
        self._load_optim_states(optimizer, optimizer.model.unwrap(), by_name)
        if "param_groups" in sd:
            self._load_param_group_hparams(optimizer.unwrap(), sd["param_groups"])
