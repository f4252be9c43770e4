"""Plugin ABCs.  Parity: reference `colossalai/booster/plugin/{plugin_base,dp_plugin_base,pp_plugin_base}.py`."""
from __future__ import annotations

import random
from abc import ABC, abstractmethod
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler as LRScheduler
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

from ...checkpoint_io import CheckpointIO
from ...interface import OptimizerWrapper

__all__ = ["Plugin", "DPPluginBase", "PipelinePluginBase"]


class Plugin(ABC):
    @abstractmethod
    def supported_devices(self) -> List[str]:
        ...

    @abstractmethod
    def supported_precisions(self) -> List[str]:
        ...

    @abstractmethod
    def control_precision(self) -> bool:
        ...

    @abstractmethod
    def control_device(self) -> bool:
        ...

    @abstractmethod
    def support_no_sync(self) -> bool:
        ...

    @abstractmethod
    def support_lora(self) -> bool:
        ...

    @abstractmethod
    def configure(self, model: nn.Module, optimizer: Optional[Optimizer] = None, criterion: Optional[Callable] = None,
                  dataloader: Optional[DataLoader] = None, lr_scheduler: Optional[LRScheduler] = None
                  ) -> Tuple[nn.Module, OptimizerWrapper, Callable, DataLoader, LRScheduler]:
        ...

    @abstractmethod
    def control_checkpoint_io(self) -> bool:
        ...

    @abstractmethod
    def get_checkpoint_io(self) -> CheckpointIO:
        ...

    @abstractmethod
    def no_sync(self, model: nn.Module, optimizer: OptimizerWrapper) -> Iterator[None]:
        ...

    @abstractmethod
    def enable_lora(self, model: nn.Module, pretrained_dir: str, lora_config: Dict, bnb_quantization_config=None
                    ) -> nn.Module:
        ...

    @abstractmethod
    def prepare_dataloader(self, dataset: Dataset, batch_size: int, shuffle: bool = False, seed: int = 1024,
                           drop_last: bool = False, pin_memory: bool = False, num_workers: int = 0, **kwargs):
        ...


def _seed_worker(seed: int):
    def fn(worker_id):
        s = seed
        np.random.seed(s)
        torch.manual_seed(s)
        random.seed(s)

    return fn


class DPPluginBase(Plugin):
    """Pure data-parallel plugins: one replica per rank, DistributedSampler over the world."""

    def __init__(self) -> None:
        super().__init__()
        assert dist.is_initialized(), "torch.distributed is not initialised; call colossalai_b200.launch* first"
        self.rank = dist.get_rank()
        self.world_size = dist.get_world_size()

    def prepare_dataloader(self, dataset, batch_size, shuffle=False, seed=1024, drop_last=False, pin_memory=False,
                           num_workers=0, distributed_sampler_cls=None, **kwargs):
        _kwargs = kwargs.copy()
        cls = distributed_sampler_cls or DistributedSampler
        sampler = cls(dataset, num_replicas=self.world_size, rank=self.rank, shuffle=shuffle)
        return DataLoader(dataset, batch_size=batch_size, sampler=sampler, worker_init_fn=_seed_worker(seed),
                          drop_last=drop_last, pin_memory=pin_memory, num_workers=num_workers, **_kwargs)


class PipelinePluginBase(Plugin):
    @abstractmethod
    def execute_pipeline(self, data_iter: Iterator, model: nn.Module, criterion: Callable,
                         optimizer: Optional[Optimizer] = None, return_loss: bool = True,
                         return_outputs: bool = False) -> dict:
        ...
