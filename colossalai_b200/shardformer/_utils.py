"""Dotted-path attribute helpers with list indices.  Parity: reference `colossalai/shardformer/_utils.py:4-112`."""
from __future__ import annotations

import re
from typing import Any


def _tokens(key: str):
    for part in key.split("."):
        m = re.fullmatch(r"(\w+)\[(\d+)\]", part)
        if m:
            yield m.group(1)
            yield int(m.group(2))
        elif part.isdigit():
            yield int(part)
        else:
            yield part


def getattr_(obj: Any, key: str, ignore: bool = False) -> Any:
    for t in _tokens(key):
        try:
            obj = obj[t] if isinstance(t, int) else getattr(obj, t)
        except (AttributeError, IndexError, KeyError, TypeError):
            if ignore:
                return None
            raise AttributeError(f"object has no attribute path {key!r} (failed at {t!r})")
    return obj


def setattr_(obj: Any, key: str, value: Any, ignore: bool = False) -> None:
    toks = list(_tokens(key))
    for t in toks[:-1]:
        try:
            obj = obj[t] if isinstance(t, int) else getattr(obj, t)
        except (AttributeError, IndexError, KeyError, TypeError):
            if ignore:
                return
            raise AttributeError(f"object has no attribute path {key!r} (failed at {t!r})")
    last = toks[-1]
    if isinstance(last, int):
        obj[last] = value
    else:
        setattr(obj, last, value)


def hasattr_(obj: Any, key: str) -> bool:
    try:
        getattr_(obj, key)
        return True
    except AttributeError:
        return False


def set_tensors_to_none(module, exclude=None) -> None:
    """Release parameters/buffers of `module` (PP: layers not held by this stage)."""
    exclude = exclude or set()
    if module in exclude:
        return
    for child in module.children():
        set_tensors_to_none(child, exclude)
    for n in list(module._parameters.keys()):
        module._parameters[n] = None
    for n in list(module._buffers.keys()):
        module._buffers[n] = None
