"""Name -> class registries used by config-driven construction.
Parity: reference `colossalai/legacy/registry/{registry.py:1-90, __init__.py}` (`Registry.register_module`,
`get_module`, `has`; the default `LAYERS`, `MODELS`, `OPTIMIZERS`, `DATASETS`, `HOOKS`, `LOSSES`, ... registries)."""
from __future__ import annotations

from typing import Dict, List, Optional, Type

__all__ = ["Registry", "LAYERS", "MODELS", "OPTIMIZERS", "DATASETS", "DIST_GROUP_INITIALIZER", "GRADIENT_HANDLER",
           "LOSSES", "HOOKS", "TRANSFORMS", "DATA_SAMPLERS", "LR_SCHEDULERS", "SCHEDULE", "OPHOOKS"]


class Registry:
    def __init__(self, name: str, third_party_library: Optional[List] = None) -> None:
        self._name = name
        self._registry: Dict[str, Type] = {}
        self._third_party_lib = third_party_library or []

    @property
    def name(self) -> str:
        return self._name

    def register_module(self, module_class: Type) -> Type:
        """Usable as a decorator: `@LAYERS.register_module class Foo: ...`."""
        name = module_class.__name__
        assert name not in self._registry, f"{name} is already registered in {self._name}"
        self._registry[name] = module_class
        return module_class

    def get_module(self, module_name: str) -> Type:
        if module_name in self._registry:
            return self._registry[module_name]
        for lib in self._third_party_lib:
            if hasattr(lib, module_name):
                return getattr(lib, module_name)
        raise NameError(f"Module {module_name} not found in the registry {self._name}")

    def has(self, module_name: str) -> bool:
        return module_name in self._registry or any(hasattr(lib, module_name) for lib in self._third_party_lib)


import torch.nn as _nn  # noqa: E402
import torch.optim as _optim  # noqa: E402

LAYERS = Registry("layers", third_party_library=[_nn])
MODELS = Registry("models")
OPTIMIZERS = Registry("optimizers", third_party_library=[_optim])
DATASETS = Registry("datasets")
DIST_GROUP_INITIALIZER = Registry("dist_group_initializer")
GRADIENT_HANDLER = Registry("gradient_handler")
LOSSES = Registry("losses", third_party_library=[_nn])
HOOKS = Registry("hooks")
TRANSFORMS = Registry("transforms")
DATA_SAMPLERS = Registry("data_samplers")
LR_SCHEDULERS = Registry("lr_schedulers", third_party_library=[_optim.lr_scheduler])
SCHEDULE = Registry("schedules")
OPHOOKS = Registry("ophooks")
