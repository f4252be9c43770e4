"""Rank-filtered logging.  Parity: reference `colossalai/logging/logger.py:12-178`."""
from __future__ import annotations

import logging
import os
from pathlib import Path
from typing import List, Optional, Union

__all__ = ["DistributedLogger", "get_dist_logger", "disable_existing_loggers"]


def _rank() -> int:
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
    except Exception:
        pass
    return int(os.environ.get("RANK", 0))


class DistributedLogger:
    """One logger per name (singleton); each call may be restricted to `ranks=[...]`."""

    _instances = {}

    @classmethod
    def get_instance(cls, name: str = "colossalai_b200") -> "DistributedLogger":
        if name not in cls._instances:
            cls._instances[name] = cls(name)
        return cls._instances[name]

    def __init__(self, name: str):
        if name in DistributedLogger._instances:
            raise RuntimeError("use get_dist_logger() instead of constructing DistributedLogger directly")
        self._name = name
        self._logger = logging.getLogger(name)
        self._logger.propagate = False
        self._logger.setLevel(logging.INFO)
        if not self._logger.handlers:
            try:
                from rich.logging import RichHandler

                handler: logging.Handler = RichHandler(show_path=False, markup=False, rich_tracebacks=True)
                handler.setFormatter(logging.Formatter("%(name)s - %(message)s"))
            except Exception:  # pragma: no cover
                handler = logging.StreamHandler()
                handler.setFormatter(logging.Formatter("[%(asctime)s] %(levelname)s %(name)s - %(message)s"))
            self._logger.addHandler(handler)

    @staticmethod
    def _check_level(level: str) -> None:
        assert level in ("DEBUG", "INFO", "WARNING", "ERROR"), f"unknown log level {level}"

    def set_level(self, level: str) -> None:
        self._check_level(level)
        self._logger.setLevel(getattr(logging, level))

    def log_to_file(self, path: Union[str, Path], mode: str = "a", level: str = "INFO", suffix: Optional[str] = None):
        self._check_level(level)
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        fname = f"rank_{_rank()}" + (f"_{suffix}" if suffix else "") + ".log"
        fh = logging.FileHandler(path / fname, mode)
        fh.setLevel(getattr(logging, level))
        fh.setFormatter(logging.Formatter("[%(asctime)s] %(levelname)s %(name)s - %(message)s"))
        self._logger.addHandler(fh)

    def _log(self, level: str, message: str, ranks: Optional[List[int]] = None) -> None:
        if ranks is None or _rank() in ranks:
            getattr(self._logger, level)(message)

    def info(self, message: str, ranks: Optional[List[int]] = None) -> None:
        self._log("info", message, ranks)

    def warning(self, message: str, ranks: Optional[List[int]] = None) -> None:
        self._log("warning", message, ranks)

    def debug(self, message: str, ranks: Optional[List[int]] = None) -> None:
        self._log("debug", message, ranks)

    def error(self, message: str, ranks: Optional[List[int]] = None) -> None:
        self._log("error", message, ranks)


def get_dist_logger(name: str = "colossalai_b200") -> DistributedLogger:
    return DistributedLogger.get_instance(name)


def disable_existing_loggers(include: Optional[List[str]] = None, exclude: Optional[List[str]] = None) -> None:
    exclude = exclude or ["colossalai_b200"]
    for log_name in list(logging.Logger.manager.loggerDict.keys()):
        if include is not None and not any(log_name.startswith(p) for p in include):
            continue
        if any(log_name.startswith(p) for p in exclude):
            continue
        logging.getLogger(log_name).setLevel(logging.WARNING)
