"""Fused (multi-block) tensor-parallel linears: one GEMM for q|k|v or gate|up, sharded block-wise.

Parity: reference `colossalai/shardformer/layer/qkv_fused_linear.py:55-1355` (`FusedLinear1D_Col`,
`FusedLinear1D_Row`, `GPT2FusedLinearConv1D_Col/Row`, `split_fused_qkv_in_gpt2_style`,
`gather_fused_qkv_in_gpt2_style`).  A fused weight [sum(split_sizes), in] is treated as consecutive blocks;
every block is sharded along its own out-dim so each rank's local weight is [q_r | k_r | v_r].
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch
import torch.nn as nn
from torch import Tensor
from torch.distributed import ProcessGroup
from torch.nn.parameter import Parameter

from ...parallel import comm
from ...tensor.d_tensor import (
    customized_distributed_tensor_to_param,
    distribute_tensor_with_customization,
    mark_customized,
    shard_colwise,
    sharded_tensor_to_param,
)
from .linear import Linear1D_Col, Linear1D_Row, _default_weight_init
from .parallel_module import ParallelModule
from .utils import create_randomizer_with_offset

__all__ = ["FusedLinear1D_Col", "FusedLinear1D_Row", "FusedLinear", "GPT2FusedLinearConv1D_Col",
           "GPT2FusedLinearConv1D_Row", "GPT2FusedLinearConv1D", "split_fused_qkv_in_gpt2_style",
           "gather_fused_qkv_in_gpt2_style"]


def _inherit_lazy(new, native):
    """Carry the initialiser log of a lazily-built native module over to its parallel replacement."""
    from ...lazy import copy_lazy_ops

    copy_lazy_ops(getattr(native, "weight", None), getattr(new, "weight", None))
    copy_lazy_ops(getattr(native, "bias", None), getattr(new, "bias", None))
    return new


def split_fused_qkv_in_gpt2_style(qkv: Tensor, split_sizes: Sequence[int], process_group: Optional[ProcessGroup],
                                  is_transposed: bool = False) -> Tensor:
    """Global fused tensor -> this rank's [blk0_r | blk1_r | ...].  `is_transposed`: blocks live on the LAST dim
    (GPT-2 Conv1D layout [in, out]); otherwise on dim 0 ([out, in] / bias)."""
    ws, r = comm.group_size(process_group), comm.group_rank(process_group)
    dim = -1 if is_transposed else 0
    blocks = torch.split(qkv, list(split_sizes), dim=dim)
    return torch.cat([b.chunk(ws, dim=dim)[r] for b in blocks], dim=dim).contiguous()


def gather_fused_qkv_in_gpt2_style(qkv: Tensor, split_sizes: Sequence[int], process_group: Optional[ProcessGroup],
                                   is_transposed: bool = False) -> Tensor:
    """Inverse of `split_fused_qkv_in_gpt2_style` (all-gather + re-interleave blocks)."""
    ws = comm.group_size(process_group)
    if ws == 1:
        return qkv
    dim = -1 if is_transposed else 0
    base = qkv._old_detach() if hasattr(qkv, "_old_detach") else qkv.detach()
    gathered = comm.all_gather(base.contiguous().unsqueeze(0), 0, process_group)  # [ws, ...]
    local_sizes = [s // ws for s in split_sizes]
    per_rank = [torch.split(gathered[i], local_sizes, dim=dim) for i in range(ws)]
    out_blocks = [torch.cat([per_rank[i][b] for i in range(ws)], dim=dim) for b in range(len(split_sizes))]
    return torch.cat(out_blocks, dim=dim).contiguous()


class FusedLinear1D_Col(Linear1D_Col):
    """Column-parallel linear over a fused weight with `split_sizes` (e.g. [Hq*D, Hkv*D, Hkv*D])."""

    def __init__(self, in_features: int, out_features: int, split_sizes: Sequence[int], bias: bool = True,
                 dtype=None, device=None, process_group: Optional[ProcessGroup] = None, weight=None, bias_=None,
                 **kwargs) -> None:
        assert sum(split_sizes) == out_features, f"split_sizes {split_sizes} must sum to {out_features}"
        ws = comm.group_size(process_group)
        for s in split_sizes:
            assert s % ws == 0, f"every fused block ({split_sizes}) must be divisible by tp={ws}"
        super().__init__(in_features, out_features, bias=bias, dtype=dtype, device=device,
                         process_group=process_group, weight=weight, bias_=bias_, **kwargs)
        self.split_sizes = list(split_sizes)
        self.local_split_sizes = [s // ws for s in split_sizes]
        pg, ss = process_group, self.split_sizes
        gshape = (out_features, in_features)
        mark_customized(self.weight, lambda t: split_fused_qkv_in_gpt2_style(t, ss, pg, False),
                        lambda t: gather_fused_qkv_in_gpt2_style(t, ss, pg, False), gshape)
        if hasattr(self.weight, "dist_shard"):
            del self.weight.dist_shard
        self.weight.tp_shard_dim = 0          # which dim the fused blocks are split on (TP-aware optimizers need it)
        if self.bias is not None:
            mark_customized(self.bias, lambda t: split_fused_qkv_in_gpt2_style(t, ss, pg, False),
                            lambda t: gather_fused_qkv_in_gpt2_style(t, ss, pg, False), (out_features,))
            if hasattr(self.bias, "dist_shard"):
                del self.bias.dist_shard
            self.bias.tp_shard_dim = 0

    @staticmethod
    def from_native_module(module: nn.Module, process_group=None, split_sizes: Sequence[int] = None, **kwargs):
        if isinstance(process_group, (list, tuple)):
            process_group = process_group[0]
        in_f, out_f = module.in_features, module.out_features
        if split_sizes is None:
            n = kwargs.pop("num_splits", 2)
            split_sizes = [out_f // n] * n
        kwargs.pop("num_splits", None)
        if module.weight.device.type == "meta":
            return _inherit_lazy(FusedLinear1D_Col(in_f, out_f, split_sizes, bias=module.bias is not None, device="meta",
                                     dtype=module.weight.dtype, process_group=process_group, **kwargs), module)
        w = Parameter(split_fused_qkv_in_gpt2_style(module.weight.data, split_sizes, process_group, False))
        b = None
        if module.bias is not None:
            b = Parameter(split_fused_qkv_in_gpt2_style(module.bias.data, split_sizes, process_group, False))
        return FusedLinear1D_Col(in_f, out_f, split_sizes, bias=b is not None, process_group=process_group, weight=w,
                                 bias_=b, **kwargs)


class FusedLinear1D_Row(Linear1D_Row):
    """Row-parallel linear whose INPUT features are a fused multi-block vector (rare; e.g. fused gate|up in)."""

    def __init__(self, in_features: int, out_features: int, split_sizes: Sequence[int], **kwargs) -> None:
        assert sum(split_sizes) == in_features
        super().__init__(in_features, out_features, **kwargs)
        self.split_sizes = list(split_sizes)
        pg, ss = self.process_group, self.split_sizes
        mark_customized(self.weight, lambda t: split_fused_qkv_in_gpt2_style(t, ss, pg, True),
                        lambda t: gather_fused_qkv_in_gpt2_style(t, ss, pg, True), (out_features, in_features))
        if hasattr(self.weight, "dist_shard"):
            del self.weight.dist_shard
        self.weight.tp_shard_dim = 1

    @staticmethod
    def from_native_module(module: nn.Module, process_group=None, split_sizes: Sequence[int] = None, **kwargs):
        if isinstance(process_group, (list, tuple)):
            process_group = process_group[0]
        if module.weight.device.type == "meta":
            return _inherit_lazy(FusedLinear1D_Row(module.in_features, module.out_features, split_sizes,
                                     bias=module.bias is not None, device="meta", dtype=module.weight.dtype,
                                     process_group=process_group, **kwargs), module)
        w = Parameter(split_fused_qkv_in_gpt2_style(module.weight.data, split_sizes, process_group, True))
        return FusedLinear1D_Row(module.in_features, module.out_features, split_sizes, bias=module.bias is not None,
                                 process_group=process_group, weight=w, bias_=module.bias, **kwargs)


class FusedLinear(ParallelModule):
    """Non-TP fused linear (kept for API parity; plain nn.Linear math with ZBV-capable wgrad)."""

    def __init__(self, in_features, out_features, bias=True, dtype=None, device=None, weight=None, bias_=None,
                 use_zbv: bool = False, **kw):
        super().__init__()
        from .linear import LinearWithGradAccum

        self.inner = LinearWithGradAccum(in_features, out_features, bias, dtype, device, weight=weight, bias_=bias_,
                                         use_zbv=use_zbv)

    @staticmethod
    def from_native_module(module, process_group=None, **kw):
        return FusedLinear(module.in_features, module.out_features, module.bias is not None, weight=module.weight,
                           bias_=module.bias, **{k: v for k, v in kw.items() if k == "use_zbv"})

    def forward(self, x):
        return self.inner(x)


# ---- GPT-2 Conv1D ([in, out] weights).  Our models store standard [out, in] weights; these adapters accept a HF
# Conv1D module, transpose once at conversion time, and then behave like the fused linears above.
class GPT2FusedLinearConv1D_Col(FusedLinear1D_Col):
    @staticmethod
    def from_native_module(module: nn.Module, process_group=None, split_sizes: Sequence[int] = None, **kwargs):
        if isinstance(process_group, (list, tuple)):
            process_group = process_group[0]
        w_t = module.weight.data.t().contiguous()  # Conv1D: [in, out] -> [out, in]
        out_f, in_f = w_t.shape
        split_sizes = split_sizes or [out_f // 3] * 3
        w = Parameter(split_fused_qkv_in_gpt2_style(w_t, split_sizes, process_group, False))
        b = None
        if getattr(module, "bias", None) is not None:
            b = Parameter(split_fused_qkv_in_gpt2_style(module.bias.data, split_sizes, process_group, False))
        return GPT2FusedLinearConv1D_Col(in_f, out_f, split_sizes, bias=b is not None, process_group=process_group,
                                         weight=w, bias_=b, **kwargs)


class GPT2FusedLinearConv1D_Row(Linear1D_Row):
    @staticmethod
    def from_native_module(module: nn.Module, process_group=None, **kwargs):
        if isinstance(process_group, (list, tuple)):
            process_group = process_group[0]
        w_t = module.weight.data.t().contiguous()
        out_f, in_f = w_t.shape
        w = sharded_tensor_to_param(shard_colwise(w_t, process_group))
        return GPT2FusedLinearConv1D_Row(in_f, out_f, bias=getattr(module, "bias", None) is not None,
                                         process_group=process_group, weight=w, bias_=getattr(module, "bias", None),
                                         **kwargs)


class GPT2FusedLinearConv1D(FusedLinear):
    @staticmethod
    def from_native_module(module, process_group=None, **kw):
        w = Parameter(module.weight.data.t().contiguous())
        return GPT2FusedLinearConv1D(w.shape[1], w.shape[0], getattr(module, "bias", None) is not None, weight=w,
                                     bias_=getattr(module, "bias", None))
