"""Base classes for parallel modules (`from_native_module`) and vocab padding.

Parity: reference `colossalai/shardformer/layer/parallel_module.py:28-360` (ParallelModule / PaddingParallelModule:
state-dict hooks that save un-padded / gathered tensors and re-shard on load).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional, Union

import torch
import torch.nn as nn
from torch.distributed import ProcessGroup

from ...parallel import comm
from ...tensor.d_tensor import (
    distribute_tensor_with_spec,
    is_distributed_tensor,
    sharded_tensor_to_param,
    to_global,
)
from ...tensor.padded_tensor import is_padded_tensor, to_padded_tensor, to_unpadded_tensor

__all__ = ["ParallelModule", "PaddingParallelModule"]


class ParallelModule(nn.Module, ABC):
    @staticmethod
    @abstractmethod
    def from_native_module(module: nn.Module, process_group: Union[ProcessGroup, List[ProcessGroup]] = None,
                           **kwargs) -> "ParallelModule":
        """Build the parallel layer from a native (un-sharded) module, slicing its weights for this rank."""

    # ---- checkpoint protocol: by default state_dict() holds the LOCAL shard; `gather_dtensor=True` callers
    # (checkpoint IO) use `to_global` on tensors tagged by tensor.d_tensor.
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for name, param in self._parameters.items():
            if param is not None:
                p = param if keep_vars else param.detach()
                destination[prefix + name] = p
        for name, buf in self._buffers.items():
            if buf is not None and name not in self._non_persistent_buffers_set:
                destination[prefix + name] = buf if keep_vars else buf.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        """Accept either a LOCAL shard (same shape) or a GLOBAL tensor (re-sharded with the param's spec)."""
        for name, param in self._parameters.items():
            if param is None:
                continue
            key = prefix + name
            if key not in state_dict:
                if strict:
                    missing_keys.append(key)
                continue
            src = state_dict[key]
            if src.shape != param.shape:
                if is_padded_tensor(param):
                    src = to_padded_tensor(src, param._current_length, param._padding_dim)
                if is_distributed_tensor(param) and src.shape != param.shape:
                    src = distribute_tensor_with_spec(src, param)
            if src.shape != param.shape:
                error_msgs.append(f"size mismatch for {key}: checkpoint {tuple(src.shape)} vs param {tuple(param.shape)}")
                continue
            with torch.no_grad():
                param.copy_(src)
        for name, buf in self._buffers.items():
            key = prefix + name
            if buf is not None and name not in self._non_persistent_buffers_set and key in state_dict:
                with torch.no_grad():
                    buf.copy_(state_dict[key])
        if strict:
            known = {prefix + n for n in list(self._parameters) + list(self._buffers)}
            for key in state_dict.keys():
                if key.startswith(prefix) and "." not in key[len(prefix):] and key not in known:
                    unexpected_keys.append(key)


class PaddingParallelModule(ParallelModule):
    """Adds vocab padding: `weight` rows are padded up to a multiple of `make_vocab_size_divisible_by` (x tp)."""

    def __init__(self, new_num_embeddings: int, old_num_embeddings: int, weight: Optional[nn.Parameter],
                 bias_: Optional[nn.Parameter] = None) -> None:
        super().__init__()
        self.new_num_embeddings, self.old_num_embeddings = new_num_embeddings, old_num_embeddings
        self.weight = weight
        self.bias = bias_
        if self.weight is not None and new_num_embeddings != old_num_embeddings and \
                self.weight.shape[0] == old_num_embeddings:
            self.resize_embedding_weight()

    def resize_embedding_weight(self) -> None:
        w = to_padded_tensor(self.weight.data, self.new_num_embeddings, 0)
        self.weight = nn.Parameter(w, requires_grad=self.weight.requires_grad)

    def resize_embedding_bias(self) -> None:
        b = to_padded_tensor(self.bias.data, self.new_num_embeddings, 0)
        self.bias = nn.Parameter(b, requires_grad=self.bias.requires_grad)
