"""Deferred weight-gradient queue for zero-bubble pipeline schedules (B computes dX only, W pops dW GEMMs).
Parity: reference `colossalai/pipeline/weight_grad_store.py:4-42`."""
from __future__ import annotations

import queue
from typing import Callable, Dict, List


class WeightGradStore:
    enabled: bool = False
    _cache: List[Callable] = []
    _queues: Dict[int, "queue.Queue"] = {}

    @classmethod
    def put(cls, fn: Callable) -> None:
        cls._cache.append(fn)

    @classmethod
    def flush(cls, chunk: int = 0) -> None:
        cls._queues.setdefault(chunk, queue.Queue()).put(cls._cache)
        cls._cache = []

    @classmethod
    def pop(cls, chunk: int = 0) -> None:
        q = cls._queues.get(chunk)
        if q is None or q.empty():
            raise RuntimeError(f"WeightGradStore: no pending W pass for chunk {chunk}")
        for fn in q.get():
            fn()

    @classmethod
    def pending(cls, chunk: int = 0) -> int:
        q = cls._queues.get(chunk)
        return 0 if q is None else q.qsize()

    @classmethod
    def clear(cls) -> None:
        cls._cache = []
        cls._queues = {}
