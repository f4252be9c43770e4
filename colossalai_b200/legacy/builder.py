"""Config-driven object construction.  Parity: reference `colossalai/legacy/builder/builder.py:1-90`
(`build_from_config`, `build_from_registry`, `build_gradient_handler`)."""
from __future__ import annotations

import inspect
from typing import Any, Dict

from .registry import GRADIENT_HANDLER, Registry

__all__ = ["build_from_config", "build_from_registry", "build_gradient_handler"]


def build_from_config(module, config: Dict[str, Any]):
    assert inspect.isclass(module), "module must be a class"
    return module(**config)


def build_from_registry(config: Dict[str, Any], registry: Registry):
    """`config = dict(type="ClassName", **kwargs)`."""
    cfg = dict(config)
    name = cfg.pop("type")
    assert registry.has(name), f"{name} is not found in registry {registry.name}"
    try:
        return registry.get_module(name)(**cfg)
    except Exception as e:
        raise type(e)(f"failed to build {name} from {registry.name}: {e}") from e


def build_gradient_handler(config: Dict[str, Any], model, optimizer):
    cfg = dict(config)
    cfg["model"], cfg["optimizer"] = model, optimizer
    return build_from_registry(cfg, GRADIENT_HANDLER)
