"""CheckpointIO base.  Parity: reference `colossalai/checkpoint_io/checkpoint_io_base.py:18-454`
(save/load model, optimizer, lr scheduler; sharded vs un-sharded; async writers joined in `synchronize`)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from pathlib import Path
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler as LRScheduler

from ..interface import ModelWrapper
from .utils import has_index_file

__all__ = ["CheckpointIO"]


class CheckpointIO(ABC):
    """
    >>> io = GeneralCheckpointIO()
    >>> io.save_model(model, "ckpt_dir", shard=True, use_safetensors=True)
    >>> io.load_model(model, "ckpt_dir")
    """

    N_WRITE_ENTRIES = 32

    def __init__(self) -> None:
        super().__init__()
        self.pinned_state_dicts: Dict[int, Dict[str, torch.Tensor]] = {}
        self.async_writers: List = []

    def _sync_io(self) -> None:
        for w in self.async_writers:
            w.synchronize()
        self.async_writers.clear()

    def _sync_d2h(self) -> None:
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def synchronize(self) -> None:
        """Join outstanding asynchronous checkpoint writes."""
        self._sync_d2h()
        self._sync_io()

    def __del__(self) -> None:
        try:
            self._sync_io()
        except Exception:
            pass

    # ------------------------------------------------------------------ model
    def load_model(self, model: Union[nn.Module, ModelWrapper], checkpoint: str, strict: bool = True,
                   low_cpu_mem_mode: bool = True, num_threads: int = 1):
        index_exists, index_path = has_index_file(checkpoint)
        origin = model
        if index_exists:
            self.load_sharded_model(model, index_path, strict, low_cpu_mem_mode=low_cpu_mem_mode,
                                    num_threads=num_threads)
        else:
            path = Path(checkpoint)
            if path.is_dir():
                cands = [p for p in path.iterdir() if p.suffix in (".safetensors", ".bin", ".pt", ".pth")]
                assert len(cands) >= 1, f"no checkpoint file found in {checkpoint}"
                checkpoint = str(cands[0])
            self.load_unsharded_model(model, checkpoint, strict, low_cpu_mem_mode=low_cpu_mem_mode,
                                      num_threads=num_threads)
        return origin

    def save_model(self, model: Union[nn.Module, ModelWrapper], checkpoint: str, shard: bool = False,
                   gather_dtensor: bool = True, prefix: str = None, size_per_shard: int = 1024,
                   use_safetensors: bool = False, use_async: bool = False) -> None:
        self.synchronize()
        if use_async:
            use_safetensors = True
        if shard:
            self.save_sharded_model(model, checkpoint, gather_dtensor, prefix, size_per_shard, use_safetensors,
                                    use_async=use_async)
        else:
            self.save_unsharded_model(model, checkpoint, gather_dtensor, use_safetensors, use_async=use_async)

    # ------------------------------------------------------------------ optimizer
    def load_optimizer(self, optimizer: Optimizer, checkpoint: str, prefix: str = None, low_cpu_mem_mode: bool = True,
                       num_threads: int = 1):
        index_exists, index_path = has_index_file(checkpoint)
        if Path(checkpoint).is_dir() and not index_exists:
            raise ValueError(f"cannot find an index file in {checkpoint}")
        if index_exists:
            self.load_sharded_optimizer(optimizer, index_path, prefix, low_cpu_mem_mode=low_cpu_mem_mode,
                                        num_threads=num_threads)
        else:
            self.load_unsharded_optimizer(optimizer, checkpoint, low_cpu_mem_mode=low_cpu_mem_mode,
                                          num_threads=num_threads)

    def save_optimizer(self, optimizer: Optimizer, checkpoint: str, shard: bool = False, gather_dtensor=True,
                       prefix: str = None, size_per_shard: int = 1024, use_async: bool = False) -> None:
        self.synchronize()
        if shard:
            self.save_sharded_optimizer(optimizer, checkpoint, gather_dtensor, prefix, size_per_shard,
                                        use_async=use_async)
        else:
            self.save_unsharded_optimizer(optimizer, checkpoint, gather_dtensor, use_async=use_async)

    # ------------------------------------------------------------------ abstract
    @abstractmethod
    def load_sharded_model(self, model: nn.Module, index_file_path: str, strict: bool, low_cpu_mem_mode: bool = True,
                           num_threads: int = 1):
        ...

    @abstractmethod
    def load_unsharded_model(self, model: nn.Module, checkpoint: str, strict: bool, low_cpu_mem_mode: bool = True,
                             num_threads: int = 1):
        ...

    @abstractmethod
    def save_sharded_model(self, model: nn.Module, checkpoint: str, gather_dtensor: bool, prefix: Optional[str],
                           size_per_shard: int, use_safetensors: bool, use_async: bool = False):
        ...

    @abstractmethod
    def save_unsharded_model(self, model: nn.Module, checkpoint: str, gather_dtensor: bool, use_safetensors: bool,
                             use_async: bool = False):
        ...

    @abstractmethod
    def load_sharded_optimizer(self, optimizer: Optimizer, index_file_path: str, prefix: str,
                               low_cpu_mem_mode: bool = True, num_threads: int = 1):
        ...

    @abstractmethod
    def load_unsharded_optimizer(self, optimizer: Optimizer, checkpoint: Path, low_cpu_mem_mode: bool = True,
                                 num_threads: int = 1):
        ...

    @abstractmethod
    def save_sharded_optimizer(self, optimizer: Optimizer, checkpoint: Path, gather_dtensor: bool, prefix: str,
                               size_per_shard: int, use_async: bool = False):
        ...

    @abstractmethod
    def save_unsharded_optimizer(self, optimizer: Optimizer, checkpoint: Path, gather_dtensor: bool,
                                 use_async: bool = False):
        ...

    # ------------------------------------------------------------------ lr scheduler / lora
    def save_lr_scheduler(self, lr_scheduler: LRScheduler, checkpoint: str) -> None:
        torch.save(lr_scheduler.state_dict(), checkpoint)

    def load_lr_scheduler(self, lr_scheduler: LRScheduler, checkpoint: str) -> None:
        lr_scheduler.load_state_dict(torch.load(checkpoint, weights_only=False))

    def save_lora_as_pretrained(self, model: Union[nn.Module, ModelWrapper], checkpoint: str,
                                use_safetensors: bool = False, state_dict: Optional[dict] = None) -> None:
        from ..booster.lora import save_lora_adapters

        save_lora_adapters(model, checkpoint, use_safetensors, state_dict)
