"""Test harness helpers.  Parity: reference `colossalai/testing/utils.py:16-298`, `pytest_wrapper.py:10`.

`spawn` differs from the reference on purpose: it picks gloo on CPU boxes (the plumbing tier) and NCCL on GPU
boxes, always rendezvous on 127.0.0.1, and enforces a wall-clock timeout so a hung collective cannot hang CI.
"""
from __future__ import annotations

import functools
import gc
import os
import random
import re
import socket
import time
from inspect import signature
from typing import Any, Callable, List

import torch
import torch.multiprocessing as mp


def parameterize(argument: str, values: List[Any]) -> Callable:
    """Loop `values` for `argument` INSIDE one call (amortises process-group init across configs)."""

    def _wrapper(func):
        @functools.wraps(func)
        def _execute(*args, **kwargs):
            for v in values:
                func(*args, **{**kwargs, argument: v})

        return _execute

    return _wrapper


def rerun_on_exception(exception_type: type = Exception, pattern: str = None, max_try: int = 5) -> Callable:
    def _match(e: Exception) -> bool:
        return pattern is None or re.search(pattern, str(e)) is not None

    def _wrapper(func):
        @functools.wraps(func)
        def _run(*args, **kwargs):
            tries = 0
            while True:
                try:
                    return func(*args, **kwargs)
                except exception_type as e:  # noqa: PERF203
                    tries += 1
                    if max_try is not None and tries >= max_try or not _match(e):
                        raise
                    time.sleep(0.5)

        if signature(func).parameters:
            _run.__signature__ = signature(func)
        return _run

    return _wrapper


def rerun_if_address_is_in_use() -> Callable:
    return rerun_on_exception(exception_type=Exception, pattern=r".*(Address already in use|EADDRINUSE).*", max_try=5)


def skip_if_not_enough_gpus(min_gpus: int) -> Callable:
    def _wrapper(f):
        @functools.wraps(f)
        def _execute(*a, **k):
            if torch.cuda.device_count() >= min_gpus:
                return f(*a, **k)
            import pytest

            pytest.skip(f"needs {min_gpus} GPUs, found {torch.cuda.device_count()}")

        return _execute

    return _wrapper


def free_port() -> int:
    while True:
        port = random.randint(20000, 60000)
        with socket.socket() as s:
            try:
                s.bind(("127.0.0.1", port))
                return port
            except OSError:
                continue


def _spawn_entry(rank: int, func: Callable, world_size: int, port: int, kwargs: dict) -> None:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["LOCAL_RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(world_size)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // max(world_size, 1) // 2))
    func(rank, world_size, port, **kwargs)


def spawn(func: Callable, nprocs: int = 1, **kwargs) -> None:
    """Run `func(rank, world_size, port, **kwargs)` in `nprocs` fresh processes."""
    port = free_port()
    mp.spawn(_spawn_entry, args=(func, nprocs, port, kwargs), nprocs=nprocs, join=True)


def clear_cache_before_run() -> Callable:
    def _wrapper(f):
        @functools.wraps(f)
        def _clear(*a, **k):
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
                torch.cuda.synchronize()
            gc.collect()
            return f(*a, **k)

        return _clear

    return _wrapper


def run_on_environment_flag(name: str) -> Callable:
    """Only run the test when the environment variable `name` is "1"."""
    import pytest

    flag = os.environ.get(name.upper(), "0") == "1"
    return pytest.mark.skipif(not flag, reason=f"environment flag {name} is not set")


class DummyDataloader:
    def __init__(self, data_gen_fn: Callable, length: int = 10) -> None:
        self.data_gen_fn, self.length, self.step = data_gen_fn, length, 0

    def __iter__(self):
        self.step = 0
        return self

    def __next__(self):
        if self.step < self.length:
            self.step += 1
            return self.data_gen_fn()
        raise StopIteration

    def __len__(self) -> int:
        return self.length
