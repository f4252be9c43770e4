"""Hook every `nn.Module.__init__` inside a context.
Parity: reference `colossalai/utils/model/utils.py:41-112` (`InsertPostInitMethodToModuleSubClasses`, base of the legacy
ZeRO init context)."""
from __future__ import annotations

import functools
from typing import Callable, Optional

import torch

__all__ = ["substitute_init_recursively", "call_to_str", "InsertPostInitMethodToModuleSubClasses"]


def substitute_init_recursively(cls, func: Callable, visited: set) -> None:
    for sub in cls.__subclasses__():
        substitute_init_recursively(sub, func, visited)
        if sub not in visited:
            func(sub)
            visited.add(sub)


def call_to_str(base, *args, **kwargs) -> str:
    parts = [str(a) for a in args] + [f"{k}={v}" for k, v in kwargs.items()]
    return f"{base}({', '.join(parts)})"


class InsertPostInitMethodToModuleSubClasses:
    """Inside the context every `nn.Module` subclass calls `self._post_init_method(module, *args, **kwargs)` right after
    its own `__init__`; subclasses implement `_post_init_method` (and optionally `_pre/_post_context_exec`)."""

    def __init__(self, default_dtype: Optional[torch.dtype] = None) -> None:
        self._old_default_dtype = None
        self._default_dtype = default_dtype

    def __enter__(self):
        if self._default_dtype is not None:
            self._old_default_dtype = torch.get_default_dtype()
            torch.set_default_dtype(self._default_dtype)

        def preprocess_after(f):
            @functools.wraps(f)
            def wrapper(module: torch.nn.Module, *args, **kwargs):
                f(module, *args, **kwargs)
                self._post_init_method(module, *args, **kwargs)

            return wrapper

        def _enable_class(cls):
            cls._old_init = cls.__init__
            cls.__init__ = preprocess_after(cls.__init__)

        def _init_subclass(cls, **kwargs):
            cls.__init__ = preprocess_after(cls.__init__)

        substitute_init_recursively(torch.nn.modules.module.Module, _enable_class, set())
        torch.nn.modules.module.Module._old_init_subclass = torch.nn.modules.module.Module.__init_subclass__
        torch.nn.modules.module.Module.__init_subclass__ = classmethod(_init_subclass)
        self._pre_context_exec()
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        if self._default_dtype is not None:
            torch.set_default_dtype(self._old_default_dtype)

        def _disable_class(cls):
            if not hasattr(cls, "_old_init"):
                raise AttributeError(f"_old_init is not found in the {cls.__name__}, please make sure that you have "
                                     "imported {cls.__name__} before entering the context.")
            cls.__init__ = cls._old_init

        substitute_init_recursively(torch.nn.modules.module.Module, _disable_class, set())
        torch.nn.modules.module.Module.__init_subclass__ = torch.nn.modules.module.Module._old_init_subclass
        self._post_context_exec()
        return False if exc_type is not None else None

    def _post_init_method(self, module, *args, **kwargs):
        pass

    def _pre_context_exec(self):
        pass

    def _post_context_exec(self):
        pass
