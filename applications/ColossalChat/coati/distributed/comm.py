"""Rollout transport between producers and the consumer.

`RolloutBuffer` is a bounded, thread-safe FIFO of rollout dicts with a staleness filter (a rollout generated with
weights older than `max_staleness` optimizer versions is dropped on pop); `WeightMailbox` is the reverse channel
(latest-wins: a producer that is busy generating simply picks up the newest published weights when it comes back).
In one process they connect a producer THREAD and the consumer (the "zero-bubble" loop of `launch_zero_bubble`);
across processes `serialize_rollout / deserialize_rollout` pack a rollout into ONE contiguous uint8 tensor so it can
travel through `torch.distributed` point-to-point ops without pickling tensors.

Parity: reference `coati/distributed/comm.py` (ray object refs + `ray_broadcast_tensor_dict`) and the shared-buffer logic of
`coati/distributed/zero_bubble/{distributor.py, consumer.py}`."""
from __future__ import annotations

import io
import threading
import time
from collections import deque
from typing import Deque, Dict, Optional, Tuple

import torch

__all__ = ["RolloutBuffer", "WeightMailbox", "serialize_rollout", "deserialize_rollout", "merge_rollouts"]


def merge_rollouts(rollouts, pad_token_id: int = 0) -> Dict:
    """One rollout out of several producers' rollouts (same prompt length and `num_generations`; responses are right-
    padded to the longest).  List fields are concatenated, `model_version` becomes the OLDEST version (staleness is
    judged by the worst contributor) and `producer` records which rows came from whom."""
    rollouts = [r for r in rollouts if r is not None]
    assert rollouts, "no rollout to merge"
    if len(rollouts) == 1:
        return rollouts[0]
    P = rollouts[0]["prompt_len"]
    assert all(r["prompt_len"] == P for r in rollouts), "producers must pad prompts to the same length"
    width = max(r["sequences"].shape[1] for r in rollouts)
    seqs = [torch.nn.functional.pad(r["sequences"], (0, width - r["sequences"].shape[1]), value=pad_token_id)
            for r in rollouts]
    out = {"sequences": torch.cat(seqs, 0), "prompt_len": P,
           "attention_mask": torch.cat([r["attention_mask"] for r in rollouts], 0),
           "model_version": min(int(r.get("model_version", 0)) for r in rollouts),
           "producer": [i for i, r in enumerate(rollouts) for _ in range(r["sequences"].shape[0])]}
    for k in rollouts[0]:
        if k not in out and isinstance(rollouts[0][k], list):
            out[k] = [x for r in rollouts for x in r[k]]
    return out


class RolloutBuffer:
    def __init__(self, capacity: int = 4, max_staleness: int = 2) -> None:
        assert capacity >= 1
        self.capacity, self.max_staleness = capacity, max_staleness
        self._q: Deque[Dict] = deque()
        self._cv = threading.Condition()
        self.closed = False
        self.stats = {"pushed": 0, "popped": 0, "dropped_stale": 0, "producer_wait_s": 0.0, "consumer_wait_s": 0.0}

    def push(self, rollout: Dict, timeout: Optional[float] = None) -> bool:
        """Blocks while the buffer is full (back-pressure on the producer).  False if the buffer was closed."""
        t0 = time.perf_counter()
        with self._cv:
            while len(self._q) >= self.capacity and not self.closed:
                if not self._cv.wait(timeout):
                    return False
            self.stats["producer_wait_s"] += time.perf_counter() - t0
            if self.closed:
                return False
            self._q.append(rollout)
            self.stats["pushed"] += 1
            self._cv.notify_all()
            return True

    def pop(self, current_version: int, timeout: Optional[float] = None) -> Optional[Dict]:
        """Oldest rollout that is at most `max_staleness` versions behind `current_version`; None when closed/empty."""
        t0 = time.perf_counter()
        with self._cv:
            while True:
                while self._q:
                    r = self._q.popleft()
                    self._cv.notify_all()
                    if current_version - int(r.get("model_version", current_version)) <= self.max_staleness:
                        self.stats["popped"] += 1
                        self.stats["consumer_wait_s"] += time.perf_counter() - t0
                        return r
                    self.stats["dropped_stale"] += 1
                if self.closed:
                    return None
                if not self._cv.wait(timeout):
                    return None

    def close(self) -> None:
        with self._cv:
            self.closed = True
            self._cv.notify_all()

    def __len__(self) -> int:
        with self._cv:
            return len(self._q)


class WeightMailbox:
    """Latest-wins slot for (state_dict, version)."""

    def __init__(self) -> None:
        self._lock = threading.Lock()
        self._item: Optional[Tuple[Dict[str, torch.Tensor], int]] = None

    def publish(self, state_dict: Dict[str, torch.Tensor], version: int) -> None:
        snap = {k: v.detach().clone() for k, v in state_dict.items()}     # the consumer keeps training on its copy
        with self._lock:
            self._item = (snap, version)

    def take(self) -> Optional[Tuple[Dict[str, torch.Tensor], int]]:
        with self._lock:
            item, self._item = self._item, None
            return item


def serialize_rollout(rollout: Dict) -> torch.Tensor:
    """Tensors are stored raw (no pickle of tensor payloads), everything else through `torch.save`'s pickler."""
    meta, blobs, off = {}, [], 0
    for k, v in rollout.items():
        if torch.is_tensor(v):
            b = v.detach().cpu().contiguous().view(torch.uint8).reshape(-1) if v.numel() else torch.empty(0, dtype=torch.uint8)
            meta[k] = ("tensor", str(v.dtype).replace("torch.", ""), tuple(v.shape), off, b.numel())
            blobs.append(b)
            off += b.numel()
        else:
            meta[k] = ("object", v)
    head = io.BytesIO()
    torch.save(meta, head)
    hb = torch.frombuffer(bytearray(head.getvalue()), dtype=torch.uint8)
    size = torch.tensor([hb.numel()], dtype=torch.int64).view(torch.uint8)
    return torch.cat([size, hb] + blobs)


def deserialize_rollout(buf: torch.Tensor) -> Dict:
    buf = buf.cpu()
    n = int(buf[:8].view(torch.int64).item())
    meta = torch.load(io.BytesIO(buf[8:8 + n].numpy().tobytes()), weights_only=False)
    base = 8 + n
    out = {}
    for k, m in meta.items():
        if m[0] == "tensor":
            _, dt, shape, off, nb = m
            t = buf[base + off: base + off + nb].clone().view(getattr(torch, dt))
            out[k] = t.reshape(shape)
        else:
            out[k] = m[1]
    return out
