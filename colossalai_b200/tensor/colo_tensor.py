"""`ColoTensor` / `ColoParameter`: tensor subclasses whose ops trigger the param-op hooks.
Parity: reference `colossalai/tensor/colo_tensor.py:39-101`, `colo_parameter.py:46-95`."""
from __future__ import annotations

from typing import Optional

import torch
from torch.utils._pytree import tree_map

from .param_op_hook import ColoParamOpHookManager

__all__ = ["ColoTensor", "ColoParameter"]

_NO_HOOK = {"__get__", "__set__", "__delete__"}


def _unwrap(x):
    if isinstance(x, ColoTensor):
        return x.as_subclass(torch.nn.Parameter if isinstance(x, torch.nn.Parameter) else torch.Tensor)
    return x


class ColoTensor(torch.Tensor):
    """Plain-data tensor subclass; results of ops on it are wrapped back unless they are views used by autograd
    internals."""

    @staticmethod
    def __new__(cls, data: torch.Tensor) -> "ColoTensor":
        if data is None:
            data = torch.empty(0)
        return torch.Tensor._make_subclass(cls, data, data.requires_grad)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        with torch._C.DisableTorchFunctionSubclass():
            ret = func(*args, **kwargs)
        return ret

    def __deepcopy__(self, memo):
        if id(self) in memo:
            return memo[id(self)]
        out = ColoTensor(self.data.clone())
        memo[id(self)] = out
        return out


class ColoParameter(ColoTensor, torch.nn.Parameter):
    """A parameter whose every consuming op is bracketed by the active `ColoParamOpHook`s."""

    def __new__(cls, data: Optional[torch.Tensor] = None, requires_grad: bool = True) -> "ColoParameter":
        if data is None:
            data = torch.empty(0)
        return torch.Tensor._make_subclass(cls, data, requires_grad)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if ColoParamOpHookManager.has_hook() and name not in _NO_HOOK and not name.startswith("__"):
            params = []
            tree_map(lambda a: params.append(a) if isinstance(a, ColoParameter) else None, (args, kwargs))
            if params:
                with torch._C.DisableTorchFunctionSubclass():
                    new_args, new_kwargs = ColoParamOpHookManager.pre_op(params, *(args, kwargs))
                    func = ColoParamOpHookManager.rewrite_op(func)
                    ret = func(*new_args, **new_kwargs)
                    return ColoParamOpHookManager.post_op(params, ret)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)

    def __deepcopy__(self, memo):
        if id(self) in memo:
            return memo[id(self)]
        out = ColoParameter(self.data.clone(), self.requires_grad)
        memo[id(self)] = out
        return out

    def __reduce_ex__(self, proto):
        return ColoParameter, (self.data, self.requires_grad)
