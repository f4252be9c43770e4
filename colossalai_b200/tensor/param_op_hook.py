"""Hooks fired around every op that consumes a `ColoParameter`.

Parity: reference `colossalai/tensor/param_op_hook.py:9-170` (`ColoParamOpHook`, `ColoParamOpHookManager.use_hooks`,
pre/post forward + backward triggers through two autograd functions, `rewrite_op`).  Users: Gemini's on-demand chunk
fetch and the FP8 op rewriting hook.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from contextlib import contextmanager
from typing import Any, List, Tuple

import torch
from torch.utils._pytree import tree_flatten, tree_unflatten

__all__ = ["ColoParamOpHook", "ColoParamOpHookManager"]


class ColoParamOpHook(ABC):
    @abstractmethod
    def pre_forward(self, params: List[torch.Tensor]) -> None: ...

    @abstractmethod
    def post_forward(self, params: List[torch.Tensor]) -> None: ...

    @abstractmethod
    def pre_backward(self, params: List[torch.Tensor]) -> None: ...

    @abstractmethod
    def post_backward(self, params: List[torch.Tensor]) -> None: ...

    def rewrite_op(self, func) -> Any:
        return func


class ColoParamOpHookManager:
    """Process-wide stack of active hooks (context-manager scoped)."""

    hooks: Tuple[ColoParamOpHook, ...] = ()

    @staticmethod
    @contextmanager
    def use_hooks(*hooks: ColoParamOpHook):
        old = ColoParamOpHookManager.hooks
        ColoParamOpHookManager.hooks = tuple(hooks)
        try:
            yield
        finally:
            ColoParamOpHookManager.hooks = old

    @staticmethod
    def _fire(name: str, params: List[torch.Tensor]) -> None:
        for h in ColoParamOpHookManager.hooks:
            getattr(h, name)(params)

    @staticmethod
    def pre_op(params: List[torch.Tensor], *args: Any) -> list:
        ColoParamOpHookManager._fire("pre_forward", params)
        flat, spec = tree_flatten(args)
        idx = [i for i, a in enumerate(flat) if _wants_grad(a)]
        if idx:
            new = _PreFwdPostBwd.apply(params, *[flat[i] for i in idx])
            for i, n in zip(idx, new if isinstance(new, tuple) else (new,)):
                flat[i] = n
        return tree_unflatten(flat, spec)

    @staticmethod
    def post_op(params: List[torch.Tensor], arg: Any) -> Any:
        ColoParamOpHookManager._fire("post_forward", params)
        flat, spec = tree_flatten(arg)
        idx = [i for i, a in enumerate(flat) if _wants_grad(a)]
        if idx:
            new = _PostFwdPreBwd.apply(params, *[flat[i] for i in idx])
            for i, n in zip(idx, new if isinstance(new, tuple) else (new,)):
                flat[i] = n
        return tree_unflatten(flat, spec)

    @staticmethod
    def has_hook() -> bool:
        return len(ColoParamOpHookManager.hooks) > 0

    @staticmethod
    def rewrite_op(func) -> Any:
        for h in ColoParamOpHookManager.hooks:
            func = h.rewrite_op(func)
        return func


def _wants_grad(obj) -> bool:
    return torch.is_tensor(obj) and (obj.requires_grad or obj.grad_fn is not None)


class _PreFwdPostBwd(torch.autograd.Function):
    """identity in forward; fires `post_backward` when gradients flow back past the op's inputs"""

    @staticmethod
    def forward(ctx, params, *args):
        ctx.params = params
        return args if len(args) > 1 else args[0]

    @staticmethod
    def backward(ctx, *grads):
        ColoParamOpHookManager._fire("post_backward", ctx.params)
        return (None,) + grads


class _PostFwdPreBwd(torch.autograd.Function):
    """identity in forward; fires `pre_backward` right before the op's backward runs"""

    @staticmethod
    def forward(ctx, params, *args):
        ctx.params = params
        return args if len(args) > 1 else args[0]

    @staticmethod
    def backward(ctx, *grads):
        ColoParamOpHookManager._fire("pre_backward", ctx.params)
        return (None,) + grads
