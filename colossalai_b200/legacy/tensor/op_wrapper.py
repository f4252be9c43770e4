"""Registry of ColoTensor-aware operator implementations (reference `legacy/tensor/op_wrapper.py` `colo_op_impl`)."""
from typing import Callable, Dict

__all__ = ["colo_op_impl", "get_colo_op_impl"]

_COLOSSAL_OPS: Dict[Callable, Callable] = {}


def colo_op_impl(func: Callable):
    """`@colo_op_impl(torch.nn.functional.linear) def colo_linear(...)`: route `func` on ColoTensors to the decorated
    implementation (looked up by `ColoTensor.__torch_function__`)."""
    def deco(impl: Callable) -> Callable:
        _COLOSSAL_OPS[func] = impl
        return impl
    return deco


def get_colo_op_impl(func: Callable):
    return _COLOSSAL_OPS.get(func)
