"""Differentiable collectives used by the 2D / 2.5D / 3D layers (autograd derives the SUMMA backward passes)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ....parallel import comm


class BroadcastFwdReduceBwd(torch.autograd.Function):
    """fwd: every rank gets `x` of group-rank `src`;  bwd: gradients are summed onto `src` (others get zeros)."""

    @staticmethod
    def forward(ctx, x, src_local, group):
        ctx.src_local, ctx.group = src_local, group
        out = x.clone()
        dist.broadcast(out, src=dist.get_global_rank(group, src_local), group=group)
        return out

    @staticmethod
    def backward(ctx, dy):
        g = dy.contiguous().clone()
        dist.reduce(g, dst=dist.get_global_rank(ctx.group, ctx.src_local), group=ctx.group)
        if dist.get_rank(ctx.group) != ctx.src_local:
            g = torch.zeros_like(g)
        return g, None, None


class AllGatherFwdReduceScatterBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, group):
        ctx.dim, ctx.group = dim, group
        return comm.all_gather(x.contiguous(), dim, group)

    @staticmethod
    def backward(ctx, dy):
        return comm.reduce_scatter(dy.contiguous(), ctx.dim, ctx.group), None, None


class ReduceScatterFwdAllGatherBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, group):
        ctx.dim, ctx.group = dim, group
        return comm.reduce_scatter(x.contiguous(), dim, group)

    @staticmethod
    def backward(ctx, dy):
        return comm.all_gather(dy.contiguous(), ctx.dim, ctx.group), None, None


class AllReduceFwdIdentityBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        out = x.clone()
        dist.all_reduce(out, group=group)
        return out

    @staticmethod
    def backward(ctx, dy):
        return dy, None


class IdentityFwdAllReduceBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x

    @staticmethod
    def backward(ctx, dy):
        g = dy.contiguous().clone()
        dist.all_reduce(g, group=ctx.group)
        return g, None
