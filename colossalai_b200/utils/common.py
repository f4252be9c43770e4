"""Small helpers.  Parity: reference `colossalai/utils/common.py:17-110`."""
from __future__ import annotations

import functools
import os
import random
from contextlib import contextmanager
from pathlib import Path
from typing import Callable, Optional, Set

import numpy as np
import torch
import torch.nn as nn


def get_current_device() -> torch.device:
    from ..accelerator import get_accelerator

    return get_accelerator().get_current_device()


def ensure_path_exists(filename: str) -> None:
    Path(filename).parent.mkdir(parents=True, exist_ok=True)


@contextmanager
def conditional_context(context_manager, enable: bool = True):
    if enable:
        with context_manager:
            yield
    else:
        yield


def is_ddp_ignored(p) -> bool:
    return getattr(p, "_ddp_to_ignore", False)


def disposable(func: Callable) -> Callable:
    executed = False

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        nonlocal executed
        if not executed:
            executed = True
            return func(*args, **kwargs)

    return wrapper


def free_storage(data: torch.Tensor) -> None:
    """Release the storage of `data` in place (keeps the tensor object alive, e.g. for autograd bookkeeping)."""
    if data.untyped_storage().size() > 0:
        assert data.storage_offset() == 0
        data.untyped_storage().resize_(0)


def alloc_storage(data: torch.Tensor) -> None:
    if data.untyped_storage().size() == 0:
        data.untyped_storage().resize_(data.numel() * data.element_size())


def set_seed(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed % (2**32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_non_persistent_buffers_set(module: nn.Module, memo: Optional[Set[nn.Module]] = None, prefix: str = "",
                                   remove_duplicate: bool = True) -> Set[str]:
    """Names of buffers registered with persistent=False (they must not be checkpointed)."""
    if memo is None:
        memo = set()
    out: Set[str] = set()
    if module not in memo:
        if remove_duplicate:
            memo.add(module)
        out |= {prefix + ("." if prefix else "") + b for b in module._non_persistent_buffers_set}
        for name, sub in module._modules.items():
            if sub is None:
                continue
            out |= get_non_persistent_buffers_set(sub, memo, prefix + ("." if prefix else "") + name, remove_duplicate)
    return out
