"""Parallel embeddings.  Parity: reference `colossalai/shardformer/layer/embedding.py:30,168,241`
(`Embedding1D` hidden-sharded + gather, `PaddingEmbedding`, `VocabParallelEmbedding1D` mask + all-reduce)."""
from __future__ import annotations

from typing import List, Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.distributed import ProcessGroup

from ...parallel import comm
from ...tensor.d_tensor import mark_sharded, shard_colwise, shard_rowwise, sharded_tensor_to_param
from ...tensor.padded_tensor import to_padded_tensor
from ._operation import gather_forward_split_backward, reduce_forward, reducescatter_forward_gather_backward
from .parallel_module import PaddingParallelModule, ParallelModule
from .utils import create_randomizer_with_offset

__all__ = ["Embedding1D", "VocabParallelEmbedding1D", "PaddingEmbedding"]


def _inherit_lazy(new, native):
    """Carry the initialiser log of a lazily-built native module over to its parallel replacement."""
    from ...lazy import copy_lazy_ops

    copy_lazy_ops(getattr(native, "weight", None), getattr(new, "weight", None))
    copy_lazy_ops(getattr(native, "bias", None), getattr(new, "bias", None))
    return new


def _padded(n: int, div: int) -> int:
    return ((n + div - 1) // div) * div


class Embedding1D(ParallelModule):
    """Embedding sharded along the hidden dim; output gathered (or kept sharded with gather_output=False)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None, dtype=None,
                 device=None, process_group: Optional[ProcessGroup] = None, gather_output: bool = True,
                 weight: Optional[nn.Parameter] = None, fp8_communication: bool = False, init_std: float = 0.02,
                 **kwargs) -> None:
        super().__init__()
        self.num_embeddings, self.embed_dim = num_embeddings, embedding_dim
        self.process_group, self.padding_idx, self.gather_output = process_group, padding_idx, gather_output
        tp = comm.group_size(process_group)
        assert embedding_dim % tp == 0
        self.embed_dim_per_partition = embedding_dim // tp
        self.embed_kwargs = kwargs
        if weight is None:
            w = torch.empty(num_embeddings, self.embed_dim_per_partition, dtype=dtype, device=device)
            if w.device.type != "meta":
                with create_randomizer_with_offset(torch.initial_seed() % (2**31), process_group).fork_rng(True):
                    nn.init.normal_(w, std=init_std)
            self.weight = sharded_tensor_to_param(mark_sharded(w, 1, process_group))
        else:
            self.weight = weight

    @staticmethod
    def from_native_module(module: nn.Embedding, process_group=None, **kwargs) -> "Embedding1D":
        if isinstance(process_group, (list, tuple)):
            process_group = process_group[0]
        if module.weight.device.type == "meta":
            return _inherit_lazy(Embedding1D(module.num_embeddings, module.embedding_dim, module.padding_idx, device="meta",
                               dtype=module.weight.dtype, process_group=process_group, **kwargs), module)
        w = sharded_tensor_to_param(shard_colwise(module.weight.data, process_group))
        return Embedding1D(module.num_embeddings, module.embedding_dim, module.padding_idx,
                           process_group=process_group, weight=w, **kwargs)

    def forward(self, ids: Tensor) -> Tensor:
        out = F.embedding(ids, self.weight, self.padding_idx)
        return gather_forward_split_backward(out, -1, self.process_group) if self.gather_output else out


class PaddingEmbedding(PaddingParallelModule):
    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None, dtype=None,
                 device=None, weight: Optional[nn.Parameter] = None, make_vocab_size_divisible_by: int = 64,
                 init_std: float = 0.02, **kwargs) -> None:
        new_n = _padded(num_embeddings, make_vocab_size_divisible_by)
        if weight is None:
            weight = nn.Parameter(torch.empty(new_n, embedding_dim, dtype=dtype, device=device))
            if weight.device.type != "meta":
                nn.init.normal_(weight, std=init_std)
        super().__init__(new_n, num_embeddings, weight)
        self.num_embeddings, self.embedding_dim, self.padding_idx = new_n, embedding_dim, padding_idx

    @staticmethod
    def from_native_module(module: nn.Embedding, process_group=None, **kwargs) -> "PaddingEmbedding":
        return PaddingEmbedding(module.num_embeddings, module.embedding_dim, module.padding_idx,
                                dtype=module.weight.dtype, device=module.weight.device, weight=module.weight, **kwargs)

    def forward(self, ids: Tensor) -> Tensor:
        return F.embedding(ids, self.weight, self.padding_idx)


class VocabParallelEmbedding1D(ParallelModule):
    """Embedding sharded over the (padded) vocab: each rank looks up the ids it owns, zeros the rest, and the
    partial results are all-reduced over TP (or reduce-scattered straight into the SP layout: `sp_scatter_dim`)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None, dtype=None,
                 device=None, process_group: Optional[ProcessGroup] = None, weight: Optional[nn.Parameter] = None,
                 make_vocab_size_divisible_by: int = 64, fp8_communication: bool = False, init_std: float = 0.02,
                 sp_scatter_dim: Optional[int] = None, **kwargs) -> None:
        super().__init__()
        self.process_group = process_group
        tp, rank = comm.group_size(process_group), comm.group_rank(process_group)
        self.old_num_embeddings = num_embeddings
        self.num_embeddings = _padded(num_embeddings, make_vocab_size_divisible_by * tp)
        self.embedding_dim, self.padding_idx = embedding_dim, padding_idx
        self.num_embeddings_per_partition = self.num_embeddings // tp
        self.vocab_start_index = rank * self.num_embeddings_per_partition
        self.vocab_end_index = self.vocab_start_index + self.num_embeddings_per_partition
        self.sp_scatter_dim = sp_scatter_dim
        if weight is None:
            w = torch.empty(self.num_embeddings_per_partition, embedding_dim, dtype=dtype, device=device)
            if w.device.type != "meta":
                with create_randomizer_with_offset(torch.initial_seed() % (2**31), process_group).fork_rng(True):
                    nn.init.normal_(w, std=init_std)
                self._zero_padding(w)
            self.weight = sharded_tensor_to_param(mark_sharded(w, 0, process_group))
        else:
            self.weight = weight

    def _zero_padding(self, w: Tensor) -> None:
        with torch.no_grad():
            if self.vocab_end_index > self.old_num_embeddings:
                start = max(self.old_num_embeddings - self.vocab_start_index, 0)
                w[start:].zero_()
            if self.padding_idx is not None and self.vocab_start_index <= self.padding_idx < self.vocab_end_index:
                w[self.padding_idx - self.vocab_start_index].zero_()

    @staticmethod
    def from_native_module(module: nn.Embedding, process_group=None, **kwargs) -> "VocabParallelEmbedding1D":
        if isinstance(process_group, (list, tuple)):
            process_group = process_group[0]
        div = kwargs.get("make_vocab_size_divisible_by", 64)
        tp = comm.group_size(process_group)
        if module.weight.device.type == "meta":
            return _inherit_lazy(VocabParallelEmbedding1D(module.num_embeddings, module.embedding_dim, module.padding_idx,
                                            device="meta", dtype=module.weight.dtype, process_group=process_group,
                                            **kwargs), module)
        new_n = _padded(module.num_embeddings, div * tp)
        wp = to_padded_tensor(module.weight.data, new_n, 0)
        w = sharded_tensor_to_param(shard_rowwise(wp, process_group))
        return VocabParallelEmbedding1D(module.num_embeddings, module.embedding_dim, module.padding_idx,
                                        process_group=process_group, weight=w, **kwargs)

    def forward(self, ids: Tensor) -> Tensor:
        if comm.group_size(self.process_group) == 1:
            return F.embedding(ids, self.weight, self.padding_idx)
        mask = (ids < self.vocab_start_index) | (ids >= self.vocab_end_index)
        local = (ids - self.vocab_start_index).masked_fill(mask, 0)
        # the padding row gets no gradient (as in nn.Embedding) on the rank that owns it
        pad = self.padding_idx
        pad = pad - self.vocab_start_index if pad is not None and self.vocab_start_index <= pad < self.vocab_end_index \
            else None
        out = F.embedding(local, self.weight, pad)
        out = out.masked_fill(mask.unsqueeze(-1), 0.0)
        if self.sp_scatter_dim is not None:
            return reducescatter_forward_gather_backward(out, self.process_group, self.sp_scatter_dim)
        return reduce_forward(out, self.process_group)
