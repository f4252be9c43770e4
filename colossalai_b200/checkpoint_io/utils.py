"""Checkpoint helpers: sharding by size, (async) save of state-dict shards, HF index naming, dtensor gather.

Parity: reference `colossalai/checkpoint_io/utils.py:149-1155` (`StateDictSharder`, `save_state_dict_shards`,
`async_move_save_state_dict_shards`, `gather_distributed_param`, `load_shard_state_dict`, index-name helpers) and
`colossalai/utils/safetensors.py:162-205` (async writer glue).  The async path stages tensors in cached PINNED host
buffers (D2H on a side stream) and hands them to our native writer thread pool (`utils/aio.py`).
"""
from __future__ import annotations

import os
import re
from collections import OrderedDict
from pathlib import Path
from typing import Callable, Dict, Iterator, List, Mapping, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch.optim import Optimizer

from ..tensor.d_tensor import is_customized_distributed_tensor, is_distributed_tensor, to_global
from ..tensor.padded_tensor import is_padded_tensor, to_unpadded_tensor

SAFE_WEIGHTS_NAME = "model.safetensors"
WEIGHTS_NAME = "pytorch_model.bin"
STATES_NAME = "pytorch_optim.bin"
SAFE_STATE_NAME = "optimizer.safetensors"
SAFE_WEIGHTS_INDEX_NAME = "model.safetensors.index.json"
WEIGHTS_INDEX_NAME = "pytorch_model.bin.index.json"
STATES_INDEX_NAME = "pytorch_optim.bin.index.json"
SAFE_STATES_INDEX_NAME = "optimizer.safetensors.index.json"
GROUP_FILE_NAME = "pytorch_optim_group.bin"


def calculate_tensor_size(t: torch.Tensor) -> float:
    """MB"""
    return t.numel() * t.element_size() / 1024 / 1024


def is_safetensors_available() -> bool:
    try:
        import safetensors  # noqa: F401

        return True
    except ImportError:
        return False


def is_safetensor_checkpoint(path: str) -> bool:
    return str(path).endswith(".safetensors")


def is_dtensor_checkpoint(path: str) -> bool:
    return str(path).endswith(".*.safetensors") or str(path).endswith(".*.bin")


def gather_distributed_param(param: torch.Tensor, keep_vars: bool = False) -> torch.Tensor:
    """TP shard -> global tensor (un-padded)."""
    t = param if keep_vars else param.detach()
    if is_distributed_tensor(param) or is_customized_distributed_tensor(param):
        t = to_global(param)
    if is_padded_tensor(param):
        if not is_padded_tensor(t):
            t._padding_dim, t._origin_length, t._current_length = param._padding_dim, param._origin_length, \
                param._current_length
        t = to_unpadded_tensor(t)
    return t


class StateDictSharder:
    """Accumulate tensors until `size_per_shard` MB, then emit the block."""

    def __init__(self, size_per_shard: int) -> None:
        self.max_shard_size = size_per_shard
        self.current_block: Dict = OrderedDict()
        self.current_block_size = 0.0

    def append_param(self, name: str, tensor: torch.Tensor) -> Tuple[Optional[Dict], float]:
        size = calculate_tensor_size(tensor)
        ret, ret_size = None, 0.0
        if self.current_block_size + size > self.max_shard_size and self.current_block_size > 0:
            ret, ret_size = self.current_block, self.current_block_size
            self.current_block, self.current_block_size = OrderedDict(), 0.0
        self.current_block[name] = tensor
        self.current_block_size += size
        return ret, ret_size

    def append_optim_state(self, param_id: int, state: Dict) -> Tuple[Optional[Dict], float]:
        size = sum(calculate_tensor_size(v) for v in state.values() if torch.is_tensor(v))
        ret, ret_size = None, 0.0
        if self.current_block_size + size > self.max_shard_size and self.current_block_size > 0:
            ret, ret_size = self.current_block, self.current_block_size
            self.current_block, self.current_block_size = OrderedDict(), 0.0
        self.current_block[param_id] = state
        self.current_block_size += size
        return ret, ret_size


def get_model_base_filenames(prefix: Optional[str] = None, use_safetensors: bool = False) -> Tuple[str, str]:
    weights = SAFE_WEIGHTS_NAME if use_safetensors else WEIGHTS_NAME
    index = SAFE_WEIGHTS_INDEX_NAME if use_safetensors else WEIGHTS_INDEX_NAME
    if prefix:
        weights, index = f"{prefix}.{weights}", f"{prefix}.{index}"
    return weights, index


def get_optimizer_base_filenames(prefix: Optional[str] = None, use_safetensors: bool = False) -> Tuple[str, str, str]:
    states = SAFE_STATE_NAME if use_safetensors else STATES_NAME
    index = SAFE_STATES_INDEX_NAME if use_safetensors else STATES_INDEX_NAME
    group = GROUP_FILE_NAME
    if prefix:
        states, index, group = f"{prefix}.{states}", f"{prefix}.{index}", f"{prefix}.{group}"
    return states, index, group


def get_shard_filename(weights_name: str, idx: int) -> str:
    """pytorch_model.bin -> pytorch_model-00001.bin (the `-of-N` suffix is filled in by the index writer)."""
    return weights_name.replace(".bin", f"-{idx + 1:05d}.bin").replace(".safetensors", f"-{idx + 1:05d}.safetensors")


def has_index_file(checkpoint_path: Union[str, Path]) -> Tuple[bool, Optional[Path]]:
    p = Path(checkpoint_path)
    if p.is_file():
        if p.name.endswith(".index.json"):
            return True, p
        return False, None
    if p.is_dir():
        idx = list(p.glob("*.index.json"))
        if len(idx) == 1:
            return True, idx[0]
        if len(idx) > 1:
            pref = [i for i in idx if "optim" not in i.name]
            return True, (pref[0] if pref else idx[0])
        return False, None
    raise RuntimeError(f"invalid checkpoint path {checkpoint_path}: not a file or directory")


# ---------------------------------------------------------------------------------------- save
def _flatten_for_safetensors(sd: Mapping) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                if torch.is_tensor(vv):
                    out[f"{k}.{kk}"] = vv.contiguous()
                elif isinstance(vv, (int, float)):
                    out[f"{k}.{kk}"] = torch.tensor(vv)
        elif torch.is_tensor(v):
            out[str(k)] = v.contiguous()
    return out


def save_state_dict(state_dict: Mapping, checkpoint_file_path: str, use_safetensors: bool) -> None:
    cpu_sd = {}
    for k, v in state_dict.items():
        if torch.is_tensor(v):
            cpu_sd[k] = v.detach().cpu() if v.device.type != "cpu" else v.detach()
        else:
            cpu_sd[k] = v
    if use_safetensors:
        from safetensors.torch import save_file

        # safetensors refuses tensors that share memory: a tied / shared parameter is stored under its first name only
        # (the HuggingFace convention; loaders re-tie, `load_state_dict` of the wrappers treats the alias as present)
        seen, unique = set(), {}
        for k, v in cpu_sd.items():
            if torch.is_tensor(v) and v.numel() > 0:
                key = (v.untyped_storage().data_ptr(), v.storage_offset(), tuple(v.shape), tuple(v.stride()))
                if key in seen:
                    continue
                seen.add(key)
            unique[k] = v
        save_file(_flatten_for_safetensors(unique), checkpoint_file_path, metadata={"format": "pt"})
    else:
        torch.save(cpu_sd, checkpoint_file_path)


def save_state_dict_shards(sharded_state_dict: Iterator[Tuple[Dict, float]], checkpoint: str, index_file,
                           base_filename: str, is_master: bool, use_safetensors: bool = False,
                           use_pp_format: bool = False) -> float:
    """Write shards as they are produced; returns total size (MB)."""
    total = 0.0
    for idx, (shard, size) in enumerate(sharded_state_dict):
        if not is_master:
            continue
        shard_file = get_shard_filename(base_filename, idx)
        total += size
        for key in shard.keys():
            index_file.append_weight_map(str(key), shard_file)
        save_state_dict(shard, os.path.join(checkpoint, shard_file), use_safetensors)
    return total


def async_save_state_dict_shards(sharded_state_dict: Iterator[Tuple[Dict, float]], checkpoint: str, index_file,
                                 base_filename: str, is_master: bool, pinned_cache: Dict[str, torch.Tensor],
                                 writers: List) -> float:
    """D2H into cached pinned buffers, then background safetensors writes (joined by CheckpointIO.synchronize)."""
    from ..utils.aio import AsyncSafetensorsWriter

    total = 0.0
    for idx, (shard, size) in enumerate(sharded_state_dict):
        if not is_master:
            continue
        shard_file = get_shard_filename(base_filename, idx)
        total += size
        staged = {}
        for key, t in shard.items():
            index_file.append_weight_map(str(key), shard_file)
            t = t.detach()
            buf = pinned_cache.get(key)
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = torch.empty(t.shape, dtype=t.dtype, device="cpu",
                                  pin_memory=torch.cuda.is_available())
                pinned_cache[key] = buf
            buf.copy_(t, non_blocking=True)
            staged[str(key)] = buf
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        w = AsyncSafetensorsWriter(os.path.join(checkpoint, shard_file))
        w.write(staged)
        writers.append(w)
    return total


def shard_model_checkpoint(state_dict: Mapping, max_shard_size: int = 1024) -> Iterator[Tuple[OrderedDict, float]]:
    sharder = StateDictSharder(max_shard_size)
    for k, w in state_dict.items():
        if not torch.is_tensor(w):
            continue
        block, size = sharder.append_param(k, w)
        if block is not None:
            yield block, size
    yield sharder.current_block, sharder.current_block_size


def shard_optimizer_checkpoint(state_dict: dict, max_shard_size: int = 1024) -> Iterator[Tuple[OrderedDict, float]]:
    sharder = StateDictSharder(max_shard_size)
    for pid, st in state_dict["state"].items():
        block, size = sharder.append_optim_state(pid, st)
        if block is not None:
            yield block, size
    yield sharder.current_block, sharder.current_block_size


def save_param_groups(state_dict: dict, group_file_path: str) -> None:
    torch.save(state_dict["param_groups"], group_file_path)


def clean_folder(checkpoint_path: str, weights_name: str, shard_filenames: List[str], is_master: bool = True) -> None:
    """Remove stale shard files of a previous save with the same prefix."""
    if not is_master:
        return
    stem = weights_name.replace(".bin", "").replace(".safetensors", "")
    reg = re.compile(r"(.*?)-\d{5}-of-\d{5}")
    for fn in os.listdir(checkpoint_path):
        full = os.path.join(checkpoint_path, fn)
        if fn.startswith(stem) and os.path.isfile(full) and fn not in shard_filenames:
            base = fn.replace(".bin", "").replace(".safetensors", "")
            if reg.fullmatch(base) is not None:
                os.remove(full)


# ---------------------------------------------------------------------------------------- load
def load_state_dict(checkpoint_file_path: Union[str, Path], trusted: Optional[bool] = None) -> Dict:
    """Load one checkpoint file.  Pickle (`.bin`) files are read with `weights_only=True` (tensors and plain
    containers only - a checkpoint from an untrusted source cannot run code).  Files this framework wrote itself that
    hold richer python objects (param-group files, scheduler state) can opt in with `trusted=True` or
    `CB200_TRUSTED_CHECKPOINTS=1`."""
    p = str(checkpoint_file_path)
    if is_safetensor_checkpoint(p):
        from safetensors.torch import load_file

        return load_file(p)
    if trusted is None:
        trusted = os.environ.get("CB200_TRUSTED_CHECKPOINTS", "0") == "1"
    if trusted:
        return torch.load(p, map_location="cpu", weights_only=False)
    try:
        return torch.load(p, map_location="cpu", weights_only=True)
    except Exception as e:   # pickle.UnpicklingError on non-tensor payloads
        raise RuntimeError(
            f"{p} holds objects that `weights_only=True` refuses to unpickle ({type(e).__name__}: {e}). If the file "
            "comes from a source you trust, pass trusted=True or set CB200_TRUSTED_CHECKPOINTS=1.") from e


def load_shard_state_dict(checkpoint_file: Union[str, Path], use_safetensors: bool = False) -> Dict:
    return load_state_dict(checkpoint_file)


def unflatten_optim_state(flat: Dict[str, torch.Tensor]) -> Dict[int, Dict]:
    out: Dict[int, Dict] = {}
    for k, v in flat.items():
        pid, name = k.split(".", 1)
        out.setdefault(int(pid), {})[name] = v
    return out


def load_state_dict_into_model(model: nn.Module, state_dict: Mapping, missing_keys: List, strict: bool = False,
                               load_sub_module: bool = True) -> None:
    """Like `load_state_dict` but records missing keys across SHARDS (a key missing in this shard may be in another)."""
    sd = OrderedDict(state_dict)
    sub_missing: List[str] = []
    unexpected: List[str] = []
    errors: List[str] = []

    def load(module: nn.Module, prefix: str = "") -> None:
        module._load_from_state_dict(sd, prefix, {}, False, sub_missing, unexpected, errors)
        if load_sub_module:
            for name, child in module._modules.items():
                if child is not None:
                    load(child, prefix + name + ".")

    load(model)
    # tolerant loader records which expected keys were absent
    expected = set(model.state_dict().keys())
    present = set(sd.keys())
    missing_keys.append(sorted(expected - present))
    if errors:
        raise RuntimeError("Error(s) in loading state_dict for {}:\n\t{}".format(model.__class__.__name__,
                                                                                 "\n\t".join(errors)))


def load_param_groups_into_optimizer(optimizer: Optimizer, param_group_path: str) -> dict:
    saved_groups = torch.load(param_group_path, map_location="cpu", weights_only=False)
    if not isinstance(saved_groups, list):
        raise ValueError(f"param groups file {param_group_path} is malformed")
    groups = optimizer.param_groups
    if len(groups) != len(saved_groups):
        raise ValueError("loaded state dict has a different number of parameter groups")
    id_map = {}
    for old_g, g in zip(saved_groups, groups):
        for old_id, p in zip(old_g["params"], g["params"]):
            id_map[old_id] = p
    new_groups = []
    for old_g, g in zip(saved_groups, groups):
        ng = dict(old_g)
        ng["params"] = g["params"]
        new_groups.append(ng)
    optimizer.__dict__.update({"param_groups": new_groups})
    return id_map


def load_states_into_optimizer(optimizer: Optimizer, state_dict: dict, id_map: dict, strict: bool = False) -> None:
    def cast(param, value, key=None):
        if isinstance(value, torch.Tensor):
            if key != "step" and param.is_floating_point():
                value = value.to(param.dtype) if value.dtype != torch.float32 else value
            return value.to(param.device)
        if isinstance(value, dict):
            return {k: cast(param, v, key=k) for k, v in value.items()}
        return value

    for k, v in state_dict.items():
        if k in id_map:
            param = id_map[k]
            optimizer.state[param] = cast(param, v)
        elif not strict:
            continue
        else:
            raise KeyError(f"optimizer state for unknown param id {k}")


def sharded_optimizer_loading_epilogue(optimizer: Optimizer) -> None:
    optimizer._patch_step_function() if hasattr(optimizer, "_patch_step_function") else None
    optimizer.defaults.setdefault("differentiable", False)


def get_optimizer_state_dict_numel(state_dict: dict) -> int:
    return sum(v.numel() for st in state_dict["state"].values() for v in st.values() if torch.is_tensor(v))
