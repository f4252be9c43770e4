"""Tensor-parallel linear layers.

Parity: reference `colossalai/shardformer/layer/linear.py:40-804` (`LinearWithGradAccum`, `Linear1D_Col`,
`Linear1D_Row`, `PaddingLMHead`, `VocabParallelLMHead1D`).  Sequence-parallel modes: None | "split_gather" |
"ring" (AG->GEMM / GEMM->RS, optionally ring-decomposed) | "all_to_all" / "ring_attn" (no comm in the linears).
Activations are token-major (`[T, H]` or `[..., H]`); `seq_parallel_dim` is the sharded token/sequence dim.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor
from torch.distributed import ProcessGroup
from torch.nn.parameter import Parameter

from ...parallel import comm
from ...tensor.d_tensor import mark_sharded, shard_colwise, shard_rowwise, sharded_tensor_to_param
from ...tensor.padded_tensor import to_padded_tensor
from ._operation import (
    gather_forward_split_backward,
    linear_allreduce_forward,
    linear_gather_forward_reducescatter_backward,
    linear_reducescatter_forward_gather_backward,
    linear_with_async_comm,
    linear_with_grad_accum,
    reduce_forward,
    split_forward_gather_backward,
)
from .parallel_module import PaddingParallelModule, ParallelModule
from .utils import create_randomizer_with_offset

__all__ = ["LinearWithGradAccum", "Linear1D_Col", "Linear1D_Row", "PaddingLMHead", "VocabParallelLMHead1D"]


def _inherit_lazy(new, native):
    """Carry the initialiser log of a lazily-built native module over to its parallel replacement."""
    from ...lazy import copy_lazy_ops

    copy_lazy_ops(getattr(native, "weight", None), getattr(new, "weight", None))
    copy_lazy_ops(getattr(native, "bias", None), getattr(new, "bias", None))
    return new


def _default_weight_init(w: Tensor, std: Optional[float] = None) -> None:
    if std is not None:
        nn.init.normal_(w, mean=0.0, std=std)
    else:
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))


class LinearWithGradAccum(ParallelModule):
    """Plain linear (no TP) whose wgrad can be deferred for zero-bubble PP (`use_zbv`)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, device=None,
                 skip_bias_add: bool = False, weight: Optional[Parameter] = None, bias_: Optional[Parameter] = None,
                 use_zbv: bool = False, init_std: Optional[float] = None, **kwargs) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.skip_bias_add, self.use_zbv = skip_bias_add, use_zbv
        if skip_bias_add and not bias:
            raise ValueError("cannot skip bias add if bias is None")
        if weight is None:
            self.weight = Parameter(torch.empty(out_features, in_features, dtype=dtype, device=device))
            if self.weight.device.type != "meta":
                _default_weight_init(self.weight, init_std)
        else:
            self.weight = weight
        if bias_ is not None:
            self.bias = bias_
        elif bias:
            self.bias = Parameter(torch.zeros(out_features, dtype=dtype, device=device))
        else:
            self.bias = None

    @staticmethod
    def from_native_module(module: nn.Module, process_group=None, **kwargs) -> "LinearWithGradAccum":
        return LinearWithGradAccum(module.in_features, module.out_features, bias=module.bias is not None,
                                   weight=module.weight, bias_=module.bias, **kwargs)

    def forward(self, x: Tensor):
        bias = self.bias if not self.skip_bias_add else None
        out = linear_with_grad_accum(x, self.weight, bias, self.use_zbv)
        return (out, self.bias) if self.skip_bias_add else out


class Linear1D_Col(ParallelModule):
    """Y = X A^T + b with A sharded along its OUTPUT features: A = [A_1; ...; A_p].

    Args mirror the reference (`gather_output`, `seq_parallel_mode`, `seq_parallel_dim`, `skip_bias_add`,
    `fp8_communication`, `use_zbv`).  `weight` is stored as the local shard [out/p, in]."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, device=None,
                 process_group: Optional[ProcessGroup] = None, gather_output: bool = False,
                 seq_parallel_mode: Optional[str] = None, seq_parallel_dim: int = 0, skip_bias_add: bool = False,
                 weight: Optional[Parameter] = None, bias_: Optional[Parameter] = None,
                 fp8_communication: bool = False, use_zbv: bool = False, init_std: Optional[float] = None,
                 overlap: bool = True, **kwargs) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.gather_output, self.skip_bias_add = gather_output, skip_bias_add
        self.seq_parallel_mode, self.seq_parallel_dim = seq_parallel_mode, seq_parallel_dim
        self.process_group, self.fp8_communication, self.use_zbv, self.overlap = process_group, fp8_communication, use_zbv, overlap
        self.tp_size, self.tp_rank = comm.group_size(process_group), comm.group_rank(process_group)
        if skip_bias_add and not bias:
            raise ValueError("cannot skip bias add if bias is None")
        assert out_features % self.tp_size == 0, f"out_features {out_features} not divisible by tp {self.tp_size}"
        self.out_features_per_partition = out_features // self.tp_size
        self.randomizer = create_randomizer_with_offset(torch.initial_seed() % (2**31), process_group)
        if weight is None:
            w = torch.empty(self.out_features_per_partition, in_features, dtype=dtype, device=device)
            if w.device.type != "meta":
                with self.randomizer.fork_rng(enable_cpu=True):
                    _default_weight_init(w, init_std)
            self.weight = sharded_tensor_to_param(mark_sharded(w, 0, process_group))
        else:
            self.weight = weight
        if bias_ is not None:
            self.bias = bias_
        elif bias:
            b = torch.zeros(self.out_features_per_partition, dtype=dtype, device=device)
            self.bias = sharded_tensor_to_param(mark_sharded(b, 0, process_group))
        else:
            self.bias = None

    @staticmethod
    def from_native_module(module: nn.Module, process_group: Union[ProcessGroup, List[ProcessGroup]] = None,
                           **kwargs) -> "Linear1D_Col":
        if isinstance(process_group, (list, tuple)):
            assert len(process_group) == 1
            process_group = process_group[0]
        in_f, out_f = module.in_features, module.out_features
        tp = comm.group_size(process_group)
        if out_f % tp != 0:
            raise ValueError(f"out_features {out_f} is not divisible by tensor parallel size {tp}")
        if module.weight.device.type == "meta":
            return _inherit_lazy(Linear1D_Col(in_f, out_f, bias=module.bias is not None, device="meta", dtype=module.weight.dtype,
                                process_group=process_group, **kwargs), module)
        w = sharded_tensor_to_param(shard_rowwise(module.weight.data, process_group))
        b = sharded_tensor_to_param(shard_rowwise(module.bias.data, process_group)) if module.bias is not None else None
        return Linear1D_Col(in_f, out_f, bias=b is not None, process_group=process_group, weight=w, bias_=b, **kwargs)

    def forward(self, x: Tensor):
        bias = self.bias if not self.skip_bias_add else None
        mode = self.seq_parallel_mode
        with comm.fp8_communication(self.fp8_communication):
            if mode in ("split_gather", "ring"):
                out = linear_gather_forward_reducescatter_backward(x, self.weight, bias, self.process_group,
                                                                   self.seq_parallel_dim, ring=(mode == "ring"),
                                                                   use_zbv=self.use_zbv)
            elif mode == "pre_gathered":
                # the caller gathered the sequence once for several column linears (gather-forward /
                # reduce-scatter-backward in front of q/k/v or gate/up): dX stays a partial sum here
                out = linear_with_grad_accum(x, self.weight, bias, self.use_zbv)
            else:
                out = linear_with_async_comm(x, self.weight, bias, self.process_group, True, self.use_zbv)
        if self.gather_output:
            out = gather_forward_split_backward(out, -1, self.process_group)
        return (out, self.bias) if self.skip_bias_add else out

    def extra_repr(self) -> str:
        return f"in={self.in_features}, out={self.out_features} (local {self.out_features_per_partition}), tp={self.tp_size}, sp={self.seq_parallel_mode}"


class Linear1D_Row(ParallelModule):
    """Y = X A^T + b with A sharded along its INPUT features; output all-reduced (or reduce-scattered under SP)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, device=None,
                 process_group: Optional[ProcessGroup] = None, seq_parallel_mode: Optional[str] = None,
                 seq_parallel_dim: int = 0, parallel_input: bool = True, skip_bias_add: bool = False,
                 weight: Optional[Parameter] = None, bias_: Optional[Parameter] = None,
                 stream_chunk_num: int = 1, fp8_communication: bool = False, use_zbv: bool = False,
                 init_std: Optional[float] = None, **kwargs) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.parallel_input, self.skip_bias_add = parallel_input, skip_bias_add
        self.seq_parallel_mode, self.seq_parallel_dim = seq_parallel_mode, seq_parallel_dim
        self.process_group, self.fp8_communication, self.use_zbv = process_group, fp8_communication, use_zbv
        self.stream_chunk_num = stream_chunk_num
        self.tp_size, self.tp_rank = comm.group_size(process_group), comm.group_rank(process_group)
        if skip_bias_add and not bias:
            raise ValueError("cannot skip bias add if bias is None")
        assert in_features % self.tp_size == 0, f"in_features {in_features} not divisible by tp {self.tp_size}"
        self.in_features_per_partition = in_features // self.tp_size
        self.randomizer = create_randomizer_with_offset(torch.initial_seed() % (2**31), process_group)
        if weight is None:
            w = torch.empty(out_features, self.in_features_per_partition, dtype=dtype, device=device)
            if w.device.type != "meta":
                with self.randomizer.fork_rng(enable_cpu=True):
                    _default_weight_init(w, init_std)
            self.weight = sharded_tensor_to_param(mark_sharded(w, 1, process_group))
        else:
            self.weight = weight
        if bias_ is not None:
            self.bias = bias_
        elif bias:
            self.bias = Parameter(torch.zeros(out_features, dtype=dtype, device=device))
        else:
            self.bias = None
        if self.bias is not None and seq_parallel_mode in ("split_gather", "ring"):
            # the bias is added AFTER the reduce-scatter, i.e. on this rank's token slice only -> partial grad
            from .utils import SeqParallelUtils

            SeqParallelUtils.marked_as_sp_partial_derived_param(self.bias)

    @staticmethod
    def from_native_module(module: nn.Module, process_group: Union[ProcessGroup, List[ProcessGroup]] = None,
                           **kwargs) -> "Linear1D_Row":
        if isinstance(process_group, (list, tuple)):
            assert len(process_group) == 1
            process_group = process_group[0]
        in_f, out_f = module.in_features, module.out_features
        tp = comm.group_size(process_group)
        if in_f % tp != 0:
            raise ValueError(f"in_features {in_f} is not divisible by tensor parallel size {tp}")
        if module.weight.device.type == "meta":
            return _inherit_lazy(Linear1D_Row(in_f, out_f, bias=module.bias is not None, device="meta", dtype=module.weight.dtype,
                                process_group=process_group, **kwargs), module)
        w = sharded_tensor_to_param(shard_colwise(module.weight.data, process_group))
        return Linear1D_Row(in_f, out_f, bias=module.bias is not None, process_group=process_group, weight=w,
                            bias_=module.bias, **kwargs)

    def forward(self, x: Tensor):
        if not self.parallel_input:
            x = split_forward_gather_backward(x, -1, self.process_group)
        else:
            assert x.shape[-1] == self.weight.shape[-1], (
                f"Linear1D_Row: input last dim {x.shape[-1]} != local in_features {self.weight.shape[-1]}")
        mode = self.seq_parallel_mode
        with comm.fp8_communication(self.fp8_communication):
            if mode in ("split_gather", "ring"):
                out = linear_reducescatter_forward_gather_backward(x, self.weight, None, self.process_group,
                                                                   self.seq_parallel_dim, ring=(mode == "ring"),
                                                                   use_zbv=self.use_zbv)
            elif self.tp_size > 1:
                out = linear_allreduce_forward(x, self.weight, self.process_group, self.use_zbv)
            else:
                out = linear_with_grad_accum(x, self.weight, None, self.use_zbv)
        if self.skip_bias_add:
            return out, self.bias
        return out if self.bias is None else out + self.bias

    def extra_repr(self) -> str:
        return f"in={self.in_features} (local {self.in_features_per_partition}), out={self.out_features}, tp={self.tp_size}, sp={self.seq_parallel_mode}"


def _padded_vocab(num: int, divisor: int) -> int:
    return ((num + divisor - 1) // divisor) * divisor


class PaddingLMHead(PaddingParallelModule):
    """LM head with vocab padded to a multiple of `make_vocab_size_divisible_by` (no TP); logits sliced back."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, device=None,
                 weight: Optional[Parameter] = None, bias_: Optional[Parameter] = None,
                 make_vocab_size_divisible_by: int = 64, init_std: Optional[float] = None, **kwargs) -> None:
        new_out = _padded_vocab(out_features, make_vocab_size_divisible_by)
        if weight is None:
            weight = Parameter(torch.empty(new_out, in_features, dtype=dtype, device=device))
            if weight.device.type != "meta":
                _default_weight_init(weight, init_std)
        if bias_ is None and bias:
            bias_ = Parameter(torch.zeros(new_out, dtype=dtype, device=device))
        super().__init__(new_out, out_features, weight, bias_)
        self.in_features, self.out_features = in_features, out_features
        if self.bias is not None and self.bias.shape[0] == out_features and new_out != out_features:
            self.resize_embedding_bias()

    @staticmethod
    def from_native_module(module: nn.Module, process_group=None, **kwargs) -> "PaddingLMHead":
        return PaddingLMHead(module.in_features, module.out_features, bias=module.bias is not None,
                             dtype=module.weight.dtype, device=module.weight.device, weight=module.weight,
                             bias_=module.bias, **kwargs)

    def forward(self, x: Tensor) -> Tensor:
        out = torch.nn.functional.linear(x, self.weight, self.bias)
        return out[..., : self.old_num_embeddings]


class VocabParallelLMHead1D(Linear1D_Col):
    """LM head sharded over the (padded) vocab.  `gather_output=False` keeps logits vocab-parallel for the
    distributed cross-entropy (parallel_output)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, device=None,
                 process_group: Optional[ProcessGroup] = None, weight: Optional[Parameter] = None,
                 bias_: Optional[Parameter] = None, make_vocab_size_divisible_by: int = 64,
                 fp8_communication: bool = False, init_std: Optional[float] = None, **kwargs) -> None:
        tp = comm.group_size(process_group)
        multiple = make_vocab_size_divisible_by * tp
        new_out = _padded_vocab(out_features, multiple)
        self.old_num_embeddings_, self.new_num_embeddings_ = out_features, new_out
        Linear1D_Col.__init__(self, in_features, new_out, bias=bias, dtype=dtype, device=device,
                              process_group=process_group, weight=weight, bias_=bias_,
                              fp8_communication=fp8_communication, init_std=init_std, **kwargs)
        self.old_num_embeddings, self.new_num_embeddings = out_features, new_out
        self.out_features_unpadded = out_features

    @staticmethod
    def from_native_module(module: nn.Module, process_group: Union[ProcessGroup, List[ProcessGroup]] = None,
                           **kwargs) -> "VocabParallelLMHead1D":
        if isinstance(process_group, (list, tuple)):
            process_group = process_group[0]
        in_f, out_f = module.in_features, module.out_features
        div = kwargs.get("make_vocab_size_divisible_by", 64)
        tp = comm.group_size(process_group)
        new_out = _padded_vocab(out_f, div * tp)
        if module.weight.device.type == "meta":
            return _inherit_lazy(VocabParallelLMHead1D(in_f, out_f, bias=module.bias is not None, device="meta",
                                         dtype=module.weight.dtype, process_group=process_group, **kwargs), module)
        wp = to_padded_tensor(module.weight.data, new_out, 0)
        w = sharded_tensor_to_param(shard_rowwise(wp, process_group))
        b = None
        if module.bias is not None:
            b = sharded_tensor_to_param(shard_rowwise(to_padded_tensor(module.bias.data, new_out, 0), process_group))
        return VocabParallelLMHead1D(in_f, out_f, bias=b is not None, process_group=process_group, weight=w, bias_=b,
                                     **kwargs)

    def forward(self, x: Tensor):
        out = Linear1D_Col.forward(self, x)
        if self.gather_output:
            out = out[..., : self.old_num_embeddings]
        return out
