"""`MetaTensor`: a tensor that computes shapes only (its payload lives on the `meta` device) but still answers
`.device` with the device it pretends to be on, so device-dependent model code traces unchanged.

Parity: reference `colossalai/_analyzer/_subclasses/meta_tensor.py` (`MetaTensor`, `MetaTensorMode`).  Built on
`__torch_dispatch__` wrapper subclasses: every aten op is replayed on the unwrapped meta payloads and the results are
wrapped again, carrying the pretended device along."""
from __future__ import annotations

from typing import Optional

import torch
from torch.overrides import TorchFunctionMode
from torch.utils._pytree import tree_map

__all__ = ["MetaTensor", "MetaTensorMode"]

_FACTORIES = {torch.empty, torch.zeros, torch.ones, torch.full, torch.rand, torch.randn, torch.randint, torch.arange,
              torch.tensor, torch.eye, torch.linspace, torch.empty_like, torch.zeros_like, torch.ones_like,
              torch.rand_like, torch.randn_like, torch.full_like, torch.empty_strided}


class MetaTensor(torch.Tensor):
    _tensor: torch.Tensor

    @staticmethod
    def __new__(cls, elem: torch.Tensor, device: Optional[torch.device] = None):
        if isinstance(elem, MetaTensor):
            device = device if device is not None else elem.device
            elem = elem._tensor
        device = torch.device(device) if device is not None else (
            elem.device if elem.device.type != "meta" else torch.device("cpu"))
        payload = elem if elem.device.type == "meta" else elem.detach().to("meta")
        r = torch.Tensor._make_wrapper_subclass(cls, payload.size(), strides=payload.stride(),
                                                storage_offset=payload.storage_offset(), dtype=payload.dtype,
                                                layout=payload.layout, device=device,
                                                requires_grad=elem.requires_grad)
        r._tensor = payload
        return r

    def __repr__(self) -> str:
        return f"MetaTensor(shape={tuple(self.shape)}, dtype={self.dtype}, device='{self.device}')"

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        device = None

        def unwrap(x):
            nonlocal device
            if isinstance(x, MetaTensor):
                device = device or x.device
                return x._tensor
            if isinstance(x, torch.Tensor) and x.device.type != "meta":
                return x.to("meta")
            return x

        args = tree_map(unwrap, args)
        kwargs = tree_map(unwrap, kwargs)
        if "device" in kwargs and kwargs["device"] is not None:
            device = torch.device(kwargs["device"])          # .to(device) / factory with an explicit device
            kwargs["device"] = torch.device("meta")
        out = func(*args, **kwargs)

        def wrap(x):
            return MetaTensor(x, device=device) if isinstance(x, torch.Tensor) and not isinstance(x, MetaTensor) else x

        return tree_map(wrap, out)

    # convenience mirrors of the reference API
    def data_ptr(self) -> int:
        return self._tensor.data_ptr() if self._tensor.device.type != "meta" else 0

    def cpu(self, *a, **k):
        return MetaTensor(self._tensor, device="cpu")

    def cuda(self, device=None, *a, **k):
        return MetaTensor(self._tensor, device=f"cuda:{device}" if isinstance(device, int) else (device or "cuda:0"))


class MetaTensorMode(TorchFunctionMode):
    """Inside the mode every tensor factory allocates a `MetaTensor` (no memory), e.g. to build a 70B model on a laptop:

        with MetaTensorMode():
            model = build_model("llama3-70b")
    """

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if func in _FACTORIES:
            device = kwargs.get("device")
            if device is None or torch.device(device).type != "meta":
                kwargs["device"] = "meta"
                out = func(*args, **kwargs)
                return MetaTensor(out, device=device if device is not None else "cpu")
        return func(*args, **kwargs)
