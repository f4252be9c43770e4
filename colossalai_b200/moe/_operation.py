"""MoE collectives and gradient scalers.

Parity: reference `colossalai/moe/_operation.py:21-452` (`AllToAll`, `AllToAllUneven`, `HierarchicalAllToAll`,
`AllGather` / `ReduceScatter`, `MoeDispatch` / `MoeCombine` / `moe_cumsum`, `EPGradScalerIn/Out`,
`DPGradScalerIn/Out`).
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor
from torch.distributed import ProcessGroup

from ..parallel import comm

__all__ = ["AllGather", "ReduceScatter", "AllToAll", "AllToAllUneven", "HierarchicalAllToAll", "MoeDispatch",
           "MoeCombine", "moe_cumsum", "EPGradScalerIn", "EPGradScalerOut", "DPGradScalerIn", "DPGradScalerOut",
           "all_to_all_uneven"]


class AllGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, group: Optional[ProcessGroup] = None, overlap: bool = False):
        ctx.group = group
        return comm.all_gather(x.unsqueeze(0), 0, group)

    @staticmethod
    def backward(ctx, dy):
        return comm.reduce_scatter(dy.contiguous(), 0, ctx.group).squeeze(0), None, None


class ReduceScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, group: Optional[ProcessGroup] = None, overlap: bool = False):
        ctx.group = group
        return comm.reduce_scatter(x, 0, group).squeeze(0)

    @staticmethod
    def backward(ctx, dy):
        return comm.all_gather(dy.contiguous().unsqueeze(0), 0, ctx.group), None, None


class AllToAll(torch.autograd.Function):
    """Even all-to-all on dim 0 ([ep, ...] chunks)."""

    @staticmethod
    def forward(ctx, x: Tensor, group: Optional[ProcessGroup] = None, overlap: bool = False):
        ctx.group = group
        if comm.group_size(group) == 1:
            return x
        x = x.contiguous()
        out = torch.empty_like(x)
        dist.all_to_all_single(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, dy):
        return AllToAll.forward(ctx, dy.contiguous(), ctx.group), None, None


class AllToAllUneven(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, input_split_sizes: List[int], output_split_sizes: List[int],
                group: Optional[ProcessGroup] = None, overlap: bool = False):
        ctx.in_splits, ctx.out_splits, ctx.group = input_split_sizes, output_split_sizes, group
        return comm.all_to_all_uneven(x, input_split_sizes, output_split_sizes, group)

    @staticmethod
    def backward(ctx, dy):
        return comm.all_to_all_uneven(dy.contiguous(), ctx.out_splits, ctx.in_splits, ctx.group), None, None, None, None


def all_to_all_uneven(x: Tensor, input_split_sizes=None, output_split_sizes=None, group=None, overlap: bool = False):
    return AllToAllUneven.apply(x, input_split_sizes, output_split_sizes, group, overlap)


class HierarchicalAllToAll(torch.autograd.Function):
    """gather intra-node -> all-to-all inter-node -> scatter intra-node (multi-node EP; on one NVSwitch box the flat
    all-to-all is already optimal, so this is only used when `groups` spans nodes)."""

    @staticmethod
    def forward(ctx, x: Tensor, groups: Tuple[ProcessGroup, ProcessGroup], src_rank: int):
        if ctx is not None:
            ctx.comm_grps, ctx.src_rank = groups, src_rank
        intra, inter = groups
        local_ws = comm.group_size(intra)
        num_group = dist.get_world_size() // local_ws if dist.is_initialized() else 1
        world = local_ws * num_group
        outs = None
        if dist.get_rank() == src_rank:
            outs = [torch.empty_like(x) for _ in range(local_ws)]
        dist.gather(x, outs, dst=src_rank, group=intra)
        if dist.get_rank() == src_rank:
            t = torch.cat(outs, dim=0)
            recv = torch.empty_like(t)
            dist.all_to_all_single(recv, t.contiguous(), group=inter)
            outs = list(recv.chunk(local_ws, dim=0))
            outs = [o.contiguous() for o in outs]
        out = torch.empty_like(x)
        dist.scatter(out, outs, src=src_rank, group=intra)
        return out

    @staticmethod
    def backward(ctx, dy):
        return HierarchicalAllToAll.forward(None, dy.contiguous(), ctx.comm_grps, ctx.src_rank), None, None


# ---- capacity-based dispatch / combine (legacy top-1/top-2 MoE; kernels in kernel/csrc/moe.cu) -------------------
def moe_cumsum(inputs: Tensor, use_kernel: bool = False) -> Tensor:
    """Exclusive position of every token inside its expert's capacity buffer: cumsum(mask, 0) - 1."""
    from ..ops import moe as moe_ops

    return moe_ops.cumsum_sub_one(inputs)


class MoeDispatch(torch.autograd.Function):
    """tokens [s, h] -> expert buffer [e, c, h] using (mask [s, e], dest_idx [s])."""

    @staticmethod
    def forward(ctx, tokens: Tensor, mask: Tensor, dest_idx: Tensor, ec: int):
        from ..ops import moe as moe_ops

        s, h = tokens.shape
        out = moe_ops.dispatch_forward(s, ec, h, tokens, mask, dest_idx)
        ctx.save_for_backward(mask, dest_idx)
        ctx.s, ctx.h, ctx.ec = s, h, ec
        return out

    @staticmethod
    def backward(ctx, dy):
        from ..ops import moe as moe_ops

        mask, dest_idx = ctx.saved_tensors
        d = moe_ops.dispatch_backward(ctx.s, ctx.ec, ctx.h, dy.contiguous(), mask, dest_idx)
        return d, None, None, None


class MoeCombine(torch.autograd.Function):
    """expert buffer [e*c, h] -> tokens [s, h] = sum_k logits[s, e_k] * expert_row."""

    @staticmethod
    def forward(ctx, expert_tokens: Tensor, logits: Tensor, mask: Tensor, dest_idx: Tensor, ec: int):
        from ..ops import moe as moe_ops

        s, e = logits.shape
        h = expert_tokens.shape[-1]
        et = expert_tokens.reshape(-1, h).contiguous()
        out = moe_ops.combine_forward(s, e, ec // e, h, et, logits, mask, dest_idx)
        ctx.save_for_backward(et, logits, mask, dest_idx)
        ctx.dims = (s, e, ec, h)
        ctx.in_shape = expert_tokens.shape
        return out

    @staticmethod
    def backward(ctx, dy):
        from ..ops import moe as moe_ops

        et, logits, mask, dest_idx = ctx.saved_tensors
        s, e, ec, h = ctx.dims
        d_expert, d_logits = moe_ops.combine_backward(s, e, ec // e, h, dy.contiguous(), et, logits, mask, dest_idx)
        return d_expert.view(ctx.in_shape), d_logits, None, None, None


class _GradScaler(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, fwd_scale: float, bwd_scale: float):
        ctx.bwd_scale = bwd_scale
        return x if fwd_scale == 1.0 else x * fwd_scale

    @staticmethod
    def backward(ctx, dy):
        return (dy if ctx.bwd_scale == 1.0 else dy * ctx.bwd_scale), None, None


class EPGradScalerIn(torch.autograd.Function):
    """Scale the gradient flowing INTO the dispatch by ep_size (expert grads are averaged over fewer replicas)."""

    @staticmethod
    def forward(ctx, x: Tensor, ep_size: int):
        ctx.ep_size = ep_size
        return x

    @staticmethod
    def backward(ctx, dy):
        return (dy * ctx.ep_size if ctx.ep_size != 1 else dy), None


class EPGradScalerOut(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, ep_size: int):
        ctx.ep_size = ep_size
        return x

    @staticmethod
    def backward(ctx, dy):
        return (dy / ctx.ep_size if ctx.ep_size != 1 else dy), None


class DPGradScalerIn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, moe_dp_size: int, activated_experts: int):
        ctx.moe_dp_size, ctx.activated = moe_dp_size, activated_experts
        return x

    @staticmethod
    def backward(ctx, dy):
        if ctx.moe_dp_size != ctx.activated and ctx.activated:
            dy = dy * (ctx.moe_dp_size / ctx.activated)
        return dy, None, None


class DPGradScalerOut(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, moe_dp_size: int, activated_experts: int):
        ctx.moe_dp_size, ctx.activated = moe_dp_size, activated_experts
        return x

    @staticmethod
    def backward(ctx, dy):
        if ctx.moe_dp_size != ctx.activated and ctx.activated:
            dy = dy * (ctx.activated / ctx.moe_dp_size)
        return dy, None, None
