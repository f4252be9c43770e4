"""Legacy communication wrappers keyed by `ParallelMode`.
Parity: reference `colossalai/legacy/communication/{collective.py:1-260, p2p.py:1-420, ring.py:1-60, utils.py}`
(`all_gather`, `reduce_scatter`, `all_reduce`, `broadcast`, `reduce`, `scatter_object_list`, pipeline
`send_forward / recv_forward / send_backward / recv_backward / send_forward_recv_backward ...`, `ring_forward`)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from ...parallel import comm
from ..context import ParallelMode, global_context as gpc

__all__ = ["all_gather", "reduce_scatter", "all_reduce", "broadcast", "reduce", "scatter_object_list",
           "send_forward", "recv_forward", "send_backward", "recv_backward", "send_forward_recv_backward",
           "send_backward_recv_forward", "send_forward_recv_forward", "send_backward_recv_backward",
           "send_forward_backward_recv_forward_backward", "ring_forward", "send_obj_meta", "recv_obj_meta"]


def _group(parallel_mode: Union[ParallelMode, dist.ProcessGroup, None]):
    if isinstance(parallel_mode, ParallelMode):
        return gpc.get_group(parallel_mode)
    return parallel_mode


def all_gather(tensor: torch.Tensor, dim: int, parallel_mode, async_op: bool = False):
    g = _group(parallel_mode)
    if comm.group_size(g) == 1:
        return (tensor, None) if async_op else tensor
    out = comm.all_gather(tensor.contiguous(), dim, g)
    return (out, None) if async_op else out


def reduce_scatter(tensor: torch.Tensor, dim: int, parallel_mode, op=dist.ReduceOp.SUM, async_op: bool = False):
    g = _group(parallel_mode)
    if comm.group_size(g) == 1:
        return (tensor, None) if async_op else tensor
    out = comm.reduce_scatter(tensor.contiguous(), dim, g)
    return (out, None) if async_op else out


def all_reduce(tensor: torch.Tensor, parallel_mode, op=dist.ReduceOp.SUM, async_op: bool = False):
    g = _group(parallel_mode)
    if comm.group_size(g) == 1:
        return (tensor, None) if async_op else tensor
    out = tensor.contiguous()
    work = dist.all_reduce(out, op=op, group=g, async_op=async_op)
    return (out, work) if async_op else out


def broadcast(tensor: torch.Tensor, src: int, parallel_mode, async_op: bool = False):
    """`src` is the GLOBAL rank of the root (reference semantics)."""
    g = _group(parallel_mode)
    if comm.group_size(g) == 1:
        return (tensor, None) if async_op else tensor
    out = tensor.contiguous()
    work = dist.broadcast(out, src=src, group=g, async_op=async_op)
    return (out, work) if async_op else out


def reduce(tensor: torch.Tensor, dst: int, parallel_mode, op=dist.ReduceOp.SUM, async_op: bool = False):
    g = _group(parallel_mode)
    if comm.group_size(g) == 1:
        return (tensor, None) if async_op else tensor
    out = tensor.contiguous()
    work = dist.reduce(out, dst=dst, op=op, group=g, async_op=async_op)
    return (out, work) if async_op else out


def scatter_object_list(scatter_object_output_list: list, scatter_object_input_list: Optional[list], src: int = 0,
                        group=None) -> None:
    dist.scatter_object_list(scatter_object_output_list, scatter_object_input_list, src=src, group=_group(group))


# ----------------------------------------------------------------------------------------------- pipeline p2p
def _pipe_neighbours() -> Tuple[Optional[int], Optional[int]]:
    g = gpc.get_group(ParallelMode.PIPELINE)
    r, n = dist.get_rank(g), dist.get_world_size(g)
    prev = dist.get_global_rank(g, r - 1) if r > 0 else None
    nxt = dist.get_global_rank(g, r + 1) if r < n - 1 else None
    return prev, nxt


def send_obj_meta(obj: torch.Tensor, dst: int) -> None:
    meta = torch.tensor([obj.dim(), *obj.shape, _DTYPES.index(obj.dtype)], dtype=torch.long)
    dist.send(torch.tensor([meta.numel()], dtype=torch.long), dst)
    dist.send(meta, dst)


def recv_obj_meta(src: int) -> Tuple[torch.Size, torch.dtype]:
    n = torch.empty(1, dtype=torch.long)
    dist.recv(n, src)
    meta = torch.empty(int(n), dtype=torch.long)
    dist.recv(meta, src)
    nd = int(meta[0])
    return torch.Size(meta[1:1 + nd].tolist()), _DTYPES[int(meta[1 + nd])]


_DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.float64, torch.uint8]


def _communicate(send_next=None, send_prev=None, recv_prev_shape=None, recv_next_shape=None, dtype=torch.float32,
                 device=None):
    prev, nxt = _pipe_neighbours()
    ops, r_prev, r_next = [], None, None
    if send_prev is not None and prev is not None:
        ops.append(dist.P2POp(dist.isend, send_prev.contiguous(), prev))
    if recv_prev_shape is not None and prev is not None:
        r_prev = torch.empty(recv_prev_shape, dtype=dtype, device=device)
        ops.append(dist.P2POp(dist.irecv, r_prev, prev))
    if send_next is not None and nxt is not None:
        ops.append(dist.P2POp(dist.isend, send_next.contiguous(), nxt))
    if recv_next_shape is not None and nxt is not None:
        r_next = torch.empty(recv_next_shape, dtype=dtype, device=device)
        ops.append(dist.P2POp(dist.irecv, r_next, nxt))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return r_prev, r_next


def recv_forward(input_tensor_shape, prev_rank=None, dtype=torch.float32, device=None):
    return _communicate(recv_prev_shape=input_tensor_shape, dtype=dtype, device=device)[0]


def recv_backward(output_grad_shape, next_rank=None, dtype=torch.float32, device=None):
    return _communicate(recv_next_shape=output_grad_shape, dtype=dtype, device=device)[1]


def send_forward(output_tensor, next_rank=None) -> None:
    _communicate(send_next=output_tensor)


def send_backward(input_tensor_grad, prev_rank=None) -> None:
    _communicate(send_prev=input_tensor_grad)


def send_forward_recv_backward(output_tensor, output_grad_shape, dtype=torch.float32, device=None):
    return _communicate(send_next=output_tensor, recv_next_shape=output_grad_shape, dtype=dtype, device=device)[1]


def send_backward_recv_forward(input_tensor_grad, input_tensor_shape, dtype=torch.float32, device=None):
    return _communicate(send_prev=input_tensor_grad, recv_prev_shape=input_tensor_shape, dtype=dtype, device=device)[0]


def send_forward_recv_forward(output_tensor, input_tensor_shape, dtype=torch.float32, device=None):
    return _communicate(send_next=output_tensor, recv_prev_shape=input_tensor_shape, dtype=dtype, device=device)[0]


def send_backward_recv_backward(input_tensor_grad, output_grad_shape, dtype=torch.float32, device=None):
    return _communicate(send_prev=input_tensor_grad, recv_next_shape=output_grad_shape, dtype=dtype, device=device)[1]


def send_forward_backward_recv_forward_backward(output_tensor, input_tensor_grad, input_tensor_shape,
                                                output_grad_shape, dtype=torch.float32, device=None):
    return _communicate(send_next=output_tensor, send_prev=input_tensor_grad, recv_prev_shape=input_tensor_shape,
                        recv_next_shape=output_grad_shape, dtype=dtype, device=device)


def ring_forward(tensor_send_next: torch.Tensor, parallel_mode) -> torch.Tensor:
    """Send to the next rank of the ring, receive from the previous one (sequence-parallel RingQK / RingAV)."""
    g = _group(parallel_mode)
    n = comm.group_size(g)
    if n == 1:
        return tensor_send_next
    recv = torch.empty_like(tensor_send_next)
    for w in comm.send_recv_ring(tensor_send_next.contiguous(), recv, g):
        w.wait()
    return recv
