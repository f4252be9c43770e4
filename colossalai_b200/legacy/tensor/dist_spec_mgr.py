"""Conversion of a local tensor between distribution specs over the TP group, differentiable (reference
`legacy/tensor/dist_spec_mgr.py:28-205`): replicate <-> shard by slicing / all-gather, shard -> shard through
all-to-all when one dimension changes, gather + slice otherwise; the backward applies the inverse conversion."""
from __future__ import annotations

from contextlib import contextmanager

import torch
import torch.distributed as dist

from .distspec import DistPlacementPattern, _DistSpec
from .process_group import ProcessGroup

__all__ = ["DistSpecManager"]


def _divide(a: int, b: int) -> int:
    assert a % b == 0, f"{a} is not divisible by {b}"
    return a // b


class _Transform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, old_spec, new_spec, pg, fwd, bwd):
        ctx.old_spec, ctx.new_spec, ctx.pg, ctx.bwd = old_spec, new_spec, pg, bwd
        return fwd(tensor, old_spec, new_spec, pg)

    @staticmethod
    def backward(ctx, grad):
        return ctx.bwd(grad.contiguous(), ctx.new_spec, ctx.old_spec, ctx.pg), None, None, None, None, None


class DistSpecManager:
    _use_autograd_function: bool = True

    @staticmethod
    def _shard_as(tensor: torch.Tensor, old_spec: _DistSpec, spec: _DistSpec, pg: ProcessGroup) -> torch.Tensor:
        assert old_spec.placement == DistPlacementPattern.REPLICATE
        idx = pg.tp_local_rank()
        chunk = tensor
        # row-major numbering of the partitions over the listed dimensions
        strides = []
        total = 1
        for n in reversed(spec.num_partitions):
            strides.append(total)
            total *= n
        strides = strides[::-1]
        assert total == pg.tp_world_size(), f"{total} partitions for a TP group of {pg.tp_world_size()}"
        for d, n, s in zip(spec.dims, spec.num_partitions, strides):
            size = _divide(chunk.size(d), n)
            chunk = chunk.narrow(d, (idx // s) % n * size, size)
        return chunk.contiguous()

    @staticmethod
    def _gather(tensor: torch.Tensor, old_spec: _DistSpec, pg: ProcessGroup) -> torch.Tensor:
        assert old_spec.placement == DistPlacementPattern.SHARD
        world = pg.tp_world_size()
        if world == 1:
            return tensor
        buf = [torch.empty_like(tensor) for _ in range(world)]
        dist.all_gather(buf, tensor.contiguous(), group=pg.tp_process_group())
        # undo the row-major partition numbering, innermost dimension first
        for d, n in zip(reversed(old_spec.dims), reversed(old_spec.num_partitions)):
            buf = [torch.cat(buf[i:i + n], dim=d) for i in range(0, len(buf), n)]
        assert len(buf) == 1
        return buf[0]

    @staticmethod
    def _all_to_all(tensor: torch.Tensor, old_spec: _DistSpec, spec: _DistSpec, pg: ProcessGroup) -> torch.Tensor:
        world = pg.tp_world_size()
        if world == 1:
            return tensor
        assert len(old_spec.dims) == 1 and len(spec.dims) == 1 and old_spec.dims != spec.dims
        scatter_dim, gather_dim = spec.dims[0], old_spec.dims[0]
        pieces = [t.contiguous() for t in tensor.chunk(world, dim=scatter_dim)]
        out = [torch.empty_like(pieces[0]) for _ in range(world)]
        dist.all_to_all(out, pieces, group=pg.tp_process_group()) if dist.get_backend(pg.tp_process_group()) != "gloo" \
            else DistSpecManager._all_to_all_by_gather(out, pieces, pg)
        return torch.cat(out, dim=gather_dim).contiguous()

    @staticmethod
    def _all_to_all_by_gather(out, pieces, pg: ProcessGroup) -> None:
        """gloo has no all_to_all: every rank gathers everyone's piece list and keeps its column."""
        world, me = pg.tp_world_size(), pg.tp_local_rank()
        for src in range(world):
            buf = [torch.empty_like(pieces[0]) for _ in range(world)]
            dist.all_gather(buf, pieces[src], group=pg.tp_process_group())
            # buf[r] = piece `src` of rank r; rank `src` needs piece src of every r -> that is out on rank src
            if src == me:
                for r in range(world):
                    out[r].copy_(buf[r])

    @staticmethod
    def _r2r(tensor, old_spec, spec, pg):
        return tensor

    @staticmethod
    def _r2s(tensor, old_spec, spec, pg):
        return DistSpecManager._shard_as(tensor, old_spec, spec, pg)

    @staticmethod
    def _s2r(tensor, old_spec, spec, pg):
        return DistSpecManager._gather(tensor, old_spec, pg)

    @staticmethod
    def _s2s(tensor, old_spec, spec, pg):
        if old_spec == spec:
            return tensor
        if len(old_spec.dims) == 1 and len(spec.dims) == 1 and old_spec.dims != spec.dims:
            return DistSpecManager._all_to_all(tensor, old_spec, spec, pg)
        from .distspec import ReplicaSpec

        full = DistSpecManager._gather(tensor, old_spec, pg)
        return DistSpecManager._shard_as(full, ReplicaSpec(), spec, pg)

    @staticmethod
    def handle_trans_spec(tensor: torch.Tensor, old_spec: _DistSpec, spec: _DistSpec, pg: ProcessGroup) -> torch.Tensor:
        assert isinstance(old_spec, _DistSpec) and isinstance(spec, _DistSpec)
        table = {
            (DistPlacementPattern.REPLICATE, DistPlacementPattern.REPLICATE): (DistSpecManager._r2r, DistSpecManager._r2r),
            (DistPlacementPattern.REPLICATE, DistPlacementPattern.SHARD): (DistSpecManager._r2s, DistSpecManager._s2r),
            (DistPlacementPattern.SHARD, DistPlacementPattern.REPLICATE): (DistSpecManager._s2r, DistSpecManager._r2s),
            (DistPlacementPattern.SHARD, DistPlacementPattern.SHARD): (DistSpecManager._s2s, DistSpecManager._s2s),
        }
        fwd, bwd = table[(old_spec.placement, spec.placement)]
        if DistSpecManager._use_autograd_function and tensor.requires_grad:
            return _Transform.apply(tensor, old_spec, spec, pg, fwd, bwd)
        return fwd(tensor, old_spec, spec, pg)

    @staticmethod
    @contextmanager
    def no_grad():
        prev = DistSpecManager._use_autograd_function
        DistSpecManager._use_autograd_function = False
        try:
            yield
        finally:
            DistSpecManager._use_autograd_function = prev
