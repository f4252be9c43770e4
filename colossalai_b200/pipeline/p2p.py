"""Pipeline point-to-point communication of arbitrary pytrees.

Parity: reference `colossalai/pipeline/p2p.py:176-799` (`create_send_metadata`, `_batch_send_recv_tensor`,
`_send_recv_serialization_object`, `_communicate`, `PipelineP2PCommunication`): tensors travel through
`batch_isend_irecv`; the tree spec + non-tensor leaves + tensor shapes/dtypes travel once as pickled metadata and are
cached (`enable_metadata_cache`) while the micro-batch shape is unchanged; combined send/recv calls avoid deadlock;
optional async handles (`overlap_p2p`).  PP send/recv deliberately stays on NCCL P2P.
"""
from __future__ import annotations

import io
import pickle
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch.distributed import ProcessGroup
from torch.utils._pytree import tree_flatten, tree_unflatten

from .stage_manager import PipelineStageManager

__all__ = ["PipelineP2PCommunication", "P2PMetadata", "create_send_metadata"]

_DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.int16, torch.int8,
           torch.uint8, torch.bool, torch.float64]


def _device_for(group: Optional[ProcessGroup]) -> torch.device:
    backend = dist.get_backend(group)
    if backend == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


@dataclass
class P2PMetadata:
    tree_spec: Any
    tensor_metadata: List[Tuple[Tuple[int, ...], torch.dtype, bool]]   # (shape, dtype, requires_grad)
    non_tensor_obj_idx: List[int]
    non_tensor_objs: List[Any]


def create_send_metadata(obj: Any, strict: bool = True, return_tensor: bool = False):
    """Flatten `obj`; describe its tensors so the receiver can pre-allocate buffers."""
    leaves, spec = tree_flatten(obj)
    tensor_meta, tensors, nt_idx, nt_objs = [], [], [], []
    for i, leaf in enumerate(leaves):
        if isinstance(leaf, torch.Tensor):
            tensors.append(leaf)
            tensor_meta.append((tuple(leaf.shape), leaf.dtype, leaf.requires_grad))
        else:
            nt_idx.append(i)
            nt_objs.append(leaf)
    md = P2PMetadata(spec, tensor_meta, nt_idx, nt_objs)
    return (md, tensors) if return_tensor else md


def _send_object(obj: Any, dst: int, group: ProcessGroup) -> None:
    buf = pickle.dumps(obj)
    dev = _device_for(group)
    size = torch.tensor([len(buf)], dtype=torch.int64, device=dev)
    data = torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(dev)
    dist.send(size, dst, group=group)
    dist.send(data, dst, group=group)


def _recv_object(src: int, group: ProcessGroup) -> Any:
    dev = _device_for(group)
    size = torch.empty(1, dtype=torch.int64, device=dev)
    dist.recv(size, src, group=group)
    data = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
    dist.recv(data, src, group=group)
    return pickle.loads(data.cpu().numpy().tobytes())


class _Handles:
    def __init__(self, works, post: Optional[Callable] = None) -> None:
        self.works, self.post = works or [], post
        self.result = None

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []
        if self.post is not None:
            self.result = self.post()
            self.post = None
        return self.result


class PipelineP2PCommunication:
    def __init__(self, stage_manager: PipelineStageManager, overlap_p2p: bool = True) -> None:
        self.stage_manager = stage_manager
        self.overlap_p2p = overlap_p2p
        self.group = stage_manager.pp_group

    # ------------------------------------------------------------------ core
    def _communicate(self, obj: Any, send_dst: Optional[int], recv_src: Optional[int],
                     send_metadata: bool = True, metadata_recv: Optional[P2PMetadata] = None,
                     send_first: Optional[bool] = None):
        """Send `obj` to `send_dst` and/or receive an object from `recv_src` (world ranks).
        Returns (received object or None, handles)."""
        group = self.group
        dev = _device_for(group)
        send_tensors: List[torch.Tensor] = []
        if send_dst is not None:
            md, send_tensors = create_send_metadata(obj, return_tensor=True)
            send_tensors = [t.contiguous() if t.device == dev else t.to(dev).contiguous() for t in send_tensors]
        if send_first is None:
            send_first = True
        # ---- metadata exchange (blocking, tiny, only when not cached)
        def do_send_md():
            if send_dst is not None and send_metadata:
                _send_object(md, send_dst, group)

        def do_recv_md():
            nonlocal metadata_recv
            if recv_src is not None and metadata_recv is None:
                metadata_recv = _recv_object(recv_src, group)

        if send_first:
            do_send_md(); do_recv_md()
        else:
            do_recv_md(); do_send_md()
        # ---- tensor exchange
        ops = []
        recv_bufs: List[torch.Tensor] = []
        if recv_src is not None:
            for shape, dtype, _ in metadata_recv.tensor_metadata:
                recv_bufs.append(torch.empty(shape, dtype=dtype, device=dev))
        send_ops = [dist.P2POp(dist.isend, t, send_dst, group) for t in send_tensors] if send_dst is not None else []
        recv_ops = [dist.P2POp(dist.irecv, t, recv_src, group) for t in recv_bufs] if recv_src is not None else []
        ops = (send_ops + recv_ops) if send_first else (recv_ops + send_ops)
        works = dist.batch_isend_irecv(ops) if ops else []

        def assemble():
            if recv_src is None:
                return None
            leaves: List[Any] = [None] * (len(recv_bufs) + len(metadata_recv.non_tensor_obj_idx))
            for i, o in zip(metadata_recv.non_tensor_obj_idx, metadata_recv.non_tensor_objs):
                leaves[i] = o
            it = iter(zip(recv_bufs, metadata_recv.tensor_metadata))
            for i in range(len(leaves)):
                if i not in metadata_recv.non_tensor_obj_idx:
                    t, (_, _, rg) = next(it)
                    leaves[i] = t.requires_grad_(rg) if t.is_floating_point() else t
            return tree_unflatten(leaves, metadata_recv.tree_spec)

        handles = _Handles(works, assemble)
        handles.keepalive = send_tensors
        handles.metadata_recv = metadata_recv
        if not self.overlap_p2p:
            handles.wait()
        return handles

    # ------------------------------------------------------------------ public API (reference names)
    def recv_forward(self, prev_rank: Optional[int] = None, metadata_recv: Optional[P2PMetadata] = None):
        prev_rank = self.stage_manager.get_prev_rank() if prev_rank is None else prev_rank
        h = self._communicate(None, None, prev_rank, metadata_recv=metadata_recv)
        return (h.wait() if not self.overlap_p2p else h), h

    def recv_backward(self, next_rank: Optional[int] = None, metadata_recv: Optional[P2PMetadata] = None):
        next_rank = self.stage_manager.get_next_rank() if next_rank is None else next_rank
        h = self._communicate(None, None, next_rank, metadata_recv=metadata_recv)
        return (h.wait() if not self.overlap_p2p else h), h

    def send_forward(self, output_object: Any, next_rank: Optional[int] = None, send_metadata: bool = True):
        next_rank = self.stage_manager.get_next_rank() if next_rank is None else next_rank
        return self._communicate(output_object, next_rank, None, send_metadata=send_metadata)

    def send_backward(self, input_object: Any, prev_rank: Optional[int] = None, send_metadata: bool = True):
        prev_rank = self.stage_manager.get_prev_rank() if prev_rank is None else prev_rank
        return self._communicate(input_object, prev_rank, None, send_metadata=send_metadata)

    def send_forward_recv_backward(self, output_object: Any, next_rank: Optional[int] = None,
                                   send_metadata: bool = True, metadata_recv: Optional[P2PMetadata] = None,
                                   send_first: Optional[bool] = None):
        next_rank = self.stage_manager.get_next_rank() if next_rank is None else next_rank
        return self._communicate(output_object, next_rank, next_rank, send_metadata, metadata_recv, send_first)

    def send_backward_recv_forward(self, input_object: Any, prev_rank: Optional[int] = None,
                                   send_metadata: bool = True, metadata_recv: Optional[P2PMetadata] = None,
                                   send_first: Optional[bool] = None):
        prev_rank = self.stage_manager.get_prev_rank() if prev_rank is None else prev_rank
        return self._communicate(input_object, prev_rank, prev_rank, send_metadata, metadata_recv, send_first)

    def send_forward_recv_forward(self, output_object: Any, prev_rank: Optional[int] = None,
                                  next_rank: Optional[int] = None, send_metadata: bool = True,
                                  metadata_recv: Optional[P2PMetadata] = None, send_first: bool = True):
        prev_rank = self.stage_manager.get_prev_rank() if prev_rank is None else prev_rank
        next_rank = self.stage_manager.get_next_rank() if next_rank is None else next_rank
        return self._communicate(output_object, next_rank, prev_rank, send_metadata, metadata_recv, send_first)

    def send_backward_recv_backward(self, input_object: Any, prev_rank: Optional[int] = None,
                                    next_rank: Optional[int] = None, send_metadata: bool = True,
                                    metadata_recv: Optional[P2PMetadata] = None, send_first: bool = True):
        prev_rank = self.stage_manager.get_prev_rank() if prev_rank is None else prev_rank
        next_rank = self.stage_manager.get_next_rank() if next_rank is None else next_rank
        return self._communicate(input_object, prev_rank, next_rank, send_metadata, metadata_recv, send_first)
