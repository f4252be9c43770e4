"""Expert-parallel dispatch/combine fused with the NVLink transfer (python side of kernel/csrc/moe.cu part 3).

    rows, local_counts, ctx = dispatch(x, topk_idx)      # rows: [capacity, H] already grouped by local expert
    y = experts(rows, local_counts)                      # grouped GEMM on device-side offsets, no host sync
    out = combine(y, topk_w, ctx)                        # P2P pull + weighted sum

Both functions are autograd-aware (`FusedEPDispatch`, `FusedEPCombine`): the backward of a push is a pull and vice
versa, reusing the row positions assigned in the forward.  SURVEY §5.8 item 6; replaces the reference's
size-exchange + host sync + uneven NCCL all_to_all + re-sort (shardformer/modeling/mixtral.py:123-208).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ..kernel import loader
from ..ops._dtypes import code
from ..parallel import comm

__all__ = ["EPWorkspace", "ep_workspace", "available", "dispatch", "combine", "moe_forward_fused"]

_lib = None
ALIGN = 128     # row alignment of every local expert's segment in the receive layout (k-block of the grouped wgrad GEMM)
_workspaces: Dict[tuple, "EPWorkspace"] = {}


def _get_lib():
    global _lib
    if _lib is None:
        lib = loader.load("cb200_moe")
        lib.cb_moe_flag_words.restype = ctypes.c_int
        _lib = lib
    return _lib


class _LocalBuffer:
    """world == 1 stand-in for a symmetric buffer (plain device memory; the kernels only see pointers)."""

    def __init__(self, nbytes: int, zero: bool = False) -> None:
        dev = torch.device("cuda", torch.cuda.current_device())
        self.tensor = (torch.zeros if zero else torch.empty)(nbytes, dtype=torch.uint8, device=dev)
        self.peer_ptrs = [self.tensor.data_ptr()]

    def ptr_array(self, world: int):
        return (ctypes.c_void_p * 16)(*([self.peer_ptrs[0]] + [0] * 15))


def _make_buffer(nbytes: int, group, world: int, zero: bool = False):
    if world == 1:
        return _LocalBuffer(nbytes, zero)
    from ..parallel.fused import _SymmBuffer

    return _SymmBuffer(nbytes, group, zero=zero)


class EPWorkspace:
    """Symmetric buffers of one (group, hidden, dtype): push-target rows, pull-source rows, count matrix, flags."""

    def __init__(self, group, hidden: int, dtype: torch.dtype, num_experts: int, capacity_rows: int) -> None:
        self.group = group
        self.world = comm.group_size(group) if group is not None else 1      # None = single-rank (no EP)
        self.rank = comm.group_rank(group) if self.world > 1 else 0
        self.hidden, self.dtype, self.E = hidden, dtype, num_experts
        self.capacity = capacity_rows
        lib = _get_lib()
        esz = torch.tensor([], dtype=dtype).element_size()
        nbytes = capacity_rows * hidden * esz
        g = group if group is not None else (dist.group.WORLD if self.world > 1 else None)
        self.flags = _make_buffer(4 * lib.cb_moe_flag_words(), g, self.world, zero=True)
        self.counts = _make_buffer(4 * self.world * num_experts, g, self.world, zero=True)
        self.rows_in = _make_buffer(nbytes, g, self.world)       # dispatch target / grad-of-y target
        self.rows_out = _make_buffer(nbytes, g, self.world)      # expert outputs / grad-of-rows, pulled by peers
        dev = torch.device("cuda", torch.cuda.current_device())
        n_local = num_experts // self.world
        self.counts_local = torch.zeros(num_experts, dtype=torch.int32, device=dev)
        self.cursor = torch.zeros(num_experts, dtype=torch.int32, device=dev)
        self.done_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch = 0
        self.can_overflow = False          # set by `ep_workspace` when the capacity is below the worst case
        self._pending_meta = None
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=g)

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def view(self, buf, rows: Optional[int] = None) -> torch.Tensor:
        esz = torch.tensor([], dtype=self.dtype).element_size()
        n = self.capacity if rows is None else rows
        return buf.tensor[: n * self.hidden * esz].view(self.dtype).view(n, self.hidden)

    def check_overflow(self) -> None:
        """Raise if the last dispatch overflowed a receive buffer.  Only a workspace sized below the worst case
        (`CB200_EP_CAPACITY_FACTOR`) can overflow; for those the dispatch leaves an asynchronous copy of its flag in
        pinned memory and the COMBINE OF THE SAME LAYER reads it (the expert GEMMs sit between the two, so the wait is
        normally over) - the error surfaces before the layer's output is used, not one MoE layer later."""
        if self._pending_meta is not None:
            host, ev = self._pending_meta
            ev.synchronize()
            self._pending_meta = None
            if int(host[1]) != 0:
                raise RuntimeError(f"fused EP dispatch: receive buffer overflow (capacity {self.capacity} rows); raise "
                                   "CB200_EP_CAPACITY_FACTOR or use the nccl MoE backend")


def available(group) -> bool:
    if not torch.cuda.is_available() or os.environ.get("CB200_DISABLE_FUSED_EP", "0") == "1":
        return False
    try:
        _get_lib()
    except Exception:
        return False
    world = comm.group_size(group) if (dist.is_initialized() and group is not None) else 1
    if world == 1:
        return True
    from ..parallel import fused

    return fused.build_available() and world <= 16


def ep_workspace(group, hidden: int, dtype: torch.dtype, num_experts: int, rows_per_rank: int) -> EPWorkspace:
    world = comm.group_size(group) if (dist.is_initialized() and group is not None) else 1
    factor = float(os.environ.get("CB200_EP_CAPACITY_FACTOR", "0"))
    worst = rows_per_rank * world
    cap = worst if factor <= 0 else min(worst, int(rows_per_rank * factor))
    pad = (num_experts // world) * ALIGN                                  # every local expert's segment is ALIGN-padded
    worst_padded = (worst + 127) // 128 * 128 + pad
    cap = (cap + 127) // 128 * 128 + pad
    if world > 1:
        # symmetric allocations must have the same size on every rank: agree on the largest request
        t = torch.tensor([cap], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        cap = int(t.item())
    key = (comm.group_key(group), hidden, dtype, num_experts)
    ws = _workspaces.get(key)
    if ws is None or ws.capacity < cap:
        ws = EPWorkspace(group, hidden, dtype, num_experts, cap)
        _workspaces[key] = ws
    ws.can_overflow = ws.capacity < worst_padded
    return ws


@dataclass
class EPContext:
    ws: EPWorkspace
    topk_idx: torch.Tensor      # int32 [T, K]
    pos: torch.Tensor           # int32 [T*K] row of (t,k) in the owner's buffer
    meta: torch.Tensor          # int32 [2]
    local_counts: torch.Tensor  # int64 [n_local] rows per local expert segment, padded to ALIGN
    local_offs: torch.Tensor    # int32 [n_local] inclusive cumsum of local_counts
    real_counts: torch.Tensor   # int32 [n_local] rows actually received per local expert
    tokens: int
    K: int


def _push_assign(ws: EPWorkspace, x: torch.Tensor, idx32: torch.Tensor) -> EPContext:
    lib = _get_lib()
    T, H = x.shape
    K = idx32.shape[1]
    dev = x.device
    n_local = ws.E // ws.world
    send_off = torch.empty(ws.E, dtype=torch.int32, device=dev)
    local_counts = torch.empty(n_local, dtype=torch.int64, device=dev)
    local_offs = torch.empty(n_local, dtype=torch.int32, device=dev)
    real_counts = torch.empty(n_local, dtype=torch.int32, device=dev)
    local_counts.cb200_aligned = True       # grouped_linear: segments start / end on multiples of 128, padding is zero
    meta = torch.zeros(2, dtype=torch.int32, device=dev)
    pos = torch.empty(T * K, dtype=torch.int32, device=dev)
    ws.check_overflow()
    epoch = ws.next_epoch()
    loader.check(lib.cb_moe_ep_dispatch(
        loader.ptr(x), loader.ptr(idx32), ws.rows_in.ptr_array(ws.world), ws.flags.ptr_array(ws.world),
        ws.counts.ptr_array(ws.world), loader.ptr(ws.counts_local), loader.ptr(ws.cursor), loader.ptr(send_off),
        loader.ptr(local_counts), loader.ptr(local_offs), loader.ptr(real_counts), loader.ptr(meta), loader.ptr(pos),
        loader.ptr(ws.done_ctr), T, K, H, ctypes.c_int64(x.stride(0)), ws.E, ws.capacity, ALIGN, ws.rank, ws.world,
        ctypes.c_uint32(epoch),
        code(x.dtype), loader.stream_ptr()), "moe_ep_dispatch")
    loader.launch_counter.add("moe_ep_dispatch", 5)
    if ws.can_overflow:
        host = torch.empty(2, dtype=torch.int32, pin_memory=True)
        host.copy_(meta, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        ws._pending_meta = (host, ev)
    return EPContext(ws, idx32, pos, meta, local_counts, local_offs, real_counts, T, K)


def _take_rows(ctx: EPContext, src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Copy the rows of the local receive layout out of (or into) a symmetric buffer: live rows are copied, the padding
    of every expert segment is zero-filled, the over-allocated tail of the buffer is not touched (no traffic that
    scales with the worst-case capacity, no host sync on the row count)."""
    if dst is None:
        dst = torch.empty_like(src)
    n_local = ctx.local_offs.numel()
    loader.check(_get_lib().cb_moe_ep_take_rows(
        loader.ptr(dst), loader.ptr(src), ctypes.c_int64(src.shape[1] * src.element_size()), loader.ptr(ctx.local_offs),
        loader.ptr(ctx.real_counts), n_local, min(src.shape[0], dst.shape[0]), loader.stream_ptr()), "moe_ep_take_rows")
    loader.launch_counter.add("moe_ep_take_rows")
    return dst


def _push_known(ctx: EPContext, x: torch.Tensor, scale: Optional[torch.Tensor]) -> torch.Tensor:
    """rows_in[owner][pos[t,k]] = scale[t,k] * x[t]; returns the local receive buffer view."""
    ws = ctx.ws
    epoch = ws.next_epoch()
    loader.check(_get_lib().cb_moe_ep_push_known(
        loader.ptr(x), loader.ptr(ctx.topk_idx), loader.ptr(scale), loader.ptr(ctx.pos),
        ws.rows_in.ptr_array(ws.world), ws.flags.ptr_array(ws.world), loader.ptr(ctx.meta), loader.ptr(ws.done_ctr),
        ctx.tokens, ctx.K, x.shape[1], ctypes.c_int64(x.stride(0)), ws.E, ws.rank, ws.world, ctypes.c_uint32(epoch),
        code(x.dtype), loader.stream_ptr()), "moe_ep_push_known")
    loader.launch_counter.add("moe_ep_push", 2)
    return ws.view(ws.rows_in)


def _pull(ctx: EPContext, y_rows: torch.Tensor, w: Optional[torch.Tensor], keep: bool
          ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Publish `y_rows` (local rows in receive-buffer layout) and pull-combine this rank's tokens."""
    ws = ctx.ws
    ws.check_overflow()
    H = ws.hidden
    out_view = ws.view(ws.rows_out)
    if y_rows.data_ptr() != out_view.data_ptr():
        _take_rows(ctx, y_rows, out_view[: y_rows.shape[0]])
    out = torch.empty(ctx.tokens, H, dtype=ws.dtype, device=y_rows.device)
    ys = torch.empty(ctx.tokens * ctx.K, H, dtype=ws.dtype, device=y_rows.device) if keep else None
    epoch = ws.next_epoch()
    loader.check(_get_lib().cb_moe_ep_combine(
        loader.ptr(ctx.topk_idx), loader.ptr(w), loader.ptr(ctx.pos), ws.rows_out.ptr_array(ws.world),
        ws.flags.ptr_array(ws.world), loader.ptr(out), loader.ptr(ys), ctx.tokens, ctx.K, H, ws.E, ws.rank, ws.world,
        ctypes.c_uint32(epoch), code(ws.dtype), loader.stream_ptr()), "moe_ep_combine")
    loader.launch_counter.add("moe_ep_combine", 2)
    return out, ys


class FusedEPDispatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, ep_ctx_holder: list, ws: EPWorkspace, idx32: torch.Tensor):
        ep = _push_assign(ws, x.contiguous() if x.stride(1) != 1 else x, idx32)
        ep_ctx_holder.append(ep)
        ctx.ep = ep
        rows = _take_rows(ep, ws.view(ws.rows_in))   # autograd-owned copy (the symmetric buffer is recycled per layer)
        ctx.mark_non_differentiable(ep.local_counts)
        return rows, ep.local_counts

    @staticmethod
    def backward(ctx, d_rows, _):
        dx, _ = _pull(ctx.ep, d_rows.contiguous(), None, keep=False)
        return dx, None, None, None


class FusedEPCombine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_rows: torch.Tensor, topk_w: torch.Tensor, ep: EPContext):
        w = topk_w.reshape(-1).float().contiguous()
        out, ys = _pull(ep, y_rows, w, keep=topk_w.requires_grad)
        ctx.ep, ctx.rows = ep, y_rows.shape[0]
        ctx.w_dtype = topk_w.dtype
        ctx.save_for_backward(w, ys if ys is not None else torch.empty(0))
        return out

    @staticmethod
    def backward(ctx, dout):
        w, ys = ctx.saved_tensors
        ep = ctx.ep
        dout = dout.contiguous()
        d_rows = _take_rows(ep, _push_known(ep, dout, w)[: ctx.rows])
        dw = None
        if ys.numel():
            dw = (ys.view(ep.tokens, ep.K, -1).float() * dout.float()[:, None, :]).sum(-1).to(ctx.w_dtype)
        return d_rows, dw, None


def dispatch(x: torch.Tensor, topk_idx: torch.Tensor, num_experts: int, group) -> Tuple[torch.Tensor, torch.Tensor, EPContext]:
    T, H = x.shape
    K = topk_idx.shape[1]
    ws = ep_workspace(group, H, x.dtype, num_experts, T * K)
    holder: list = []
    rows, local_counts = FusedEPDispatch.apply(x, holder, ws, topk_idx.to(torch.int32).contiguous())
    local_counts.cb200_aligned = True
    return rows, local_counts, holder[0]


def combine(y_rows: torch.Tensor, topk_w: torch.Tensor, ep: EPContext) -> torch.Tensor:
    return FusedEPCombine.apply(y_rows, topk_w, ep)


def moe_forward_fused(x: torch.Tensor, topk_w: torch.Tensor, topk_idx: torch.Tensor, experts, num_experts: int,
                      ep_group) -> torch.Tensor:
    rows, local_counts, ep = dispatch(x, topk_idx, num_experts, ep_group)
    y = experts(rows, local_counts)
    return combine(y, topk_w, ep)
