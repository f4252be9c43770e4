"""Remember a `from_pretrained` path on a lazily-built model so `boost()` loads weights AFTER sharding.
Parity: reference `colossalai/interface/pretrained.py:11-16`."""
from typing import Optional

from torch.nn import Module

__all__ = ["get_pretrained_path", "set_pretrained_path"]


def get_pretrained_path(model: Module) -> Optional[str]:
    return getattr(model, "_pretrained", None)


def set_pretrained_path(model: Module, path: Optional[str]) -> None:
    setattr(model, "_pretrained", path)
