"""Attr-dict config + python-file configs.  Parity: reference `colossalai/context/config.py:12-107`,
`context/singleton_meta.py`."""
from __future__ import annotations

import importlib.util
import inspect
import sys
from pathlib import Path
from typing import Any, Union

__all__ = ["Config", "ConfigException", "SingletonMeta"]


class ConfigException(Exception):
    pass


class SingletonMeta(type):
    _instances: dict = {}

    def __call__(cls, *args, **kwargs):
        if cls not in cls._instances:
            cls._instances[cls] = super().__call__(*args, **kwargs)
        return cls._instances[cls]


class Config(dict):
    """dict with attribute access; nested dicts are converted recursively."""

    def __init__(self, config: dict = None):
        super().__init__()
        if config is not None:
            for k, v in config.items():
                self._add_item(k, v)

    def __missing__(self, key):
        raise KeyError(key)

    def __getattr__(self, key):
        try:
            return super().__getitem__(key)
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self._add_item(key, value)          # assigned dicts become Configs too (attribute access all the way down)

    def _add_item(self, key, value):
        self[key] = Config(value) if isinstance(value, dict) and not isinstance(value, Config) else value

    def update(self, config=None, **kw):  # type: ignore[override]
        for k, v in dict(config or {}, **kw).items():
            self._add_item(k, v)
        return self

    @staticmethod
    def from_file(filename: Union[str, Path]) -> "Config":
        filename = Path(filename)
        if not filename.exists():
            raise FileNotFoundError(filename)
        if filename.suffix != ".py":
            raise ConfigException("only .py config files are supported")
        spec = importlib.util.spec_from_file_location(f"_cb200_cfg_{filename.stem}", filename)
        module = importlib.util.module_from_spec(spec)
        assert spec.loader is not None
        spec.loader.exec_module(module)
        cfg = Config()
        for k, v in vars(module).items():
            if k.startswith("__") or inspect.ismodule(v) or inspect.isclass(v) or inspect.isfunction(v):
                continue
            cfg._add_item(k, v)
        sys.modules.pop(spec.name, None)
        return cfg
