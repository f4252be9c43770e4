"""Model wrappers.  Parity: reference `colossalai/interface/model.py:49-136`."""
from __future__ import annotations

import re
from typing import Dict, Set

import torch.nn as nn


class ModelWrapper(nn.Module):
    """Base class of every boosted model; `unwrap()` peels nested wrappers down to the user module."""

    def __init__(self, module: nn.Module) -> None:
        super().__init__()
        self.module = module

    def unwrap(self, unwrap_peft: bool = True) -> nn.Module:
        if isinstance(self.module, ModelWrapper):
            model = self.module.unwrap()
        else:
            model = self.module
        if unwrap_peft and getattr(model, "_cb200_lora_wrapped", False) and hasattr(model, "base_model"):
            model = PeftUnwrapMixin(model)
        return model

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def __getattr__(self, name: str):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)


class AMPModelMixin:
    """Hook points used by mixed-precision optimizers."""

    def update_master_params(self) -> None:
        """Called by the checkpoint IO after weights were loaded into the model: a mixed-precision / ZeRO optimizer
        registered through `bind_optimizer` re-derives its fp32 master copies from the new working weights."""
        ref = getattr(self, "_bound_optimizer", None)
        opt = ref() if ref is not None else None
        if opt is not None and hasattr(opt, "update_master_params"):
            opt.update_master_params(getattr(self, "module", self))

    def bind_optimizer(self, optimizer) -> None:
        import weakref

        object.__setattr__(self, "_bound_optimizer", weakref.ref(optimizer))


class PeftUnwrapMixin:
    """View of a LoRA-wrapped model whose state_dict names match the base model (adapters merged out of the names)."""

    def __init__(self, peft_model) -> None:
        self.base_model = getattr(peft_model, "base_model", peft_model)
        self.peft_model = peft_model

    def named_parameters(self):
        for n, p in self.peft_model.named_parameters():
            if "lora_" in n:
                continue
            yield n.replace(".base_layer", ""), p

    def state_dict(self) -> Dict:
        return {n: p for n, p in self.named_parameters()}

    def __getattr__(self, name):
        return getattr(self.peft_model, name)
