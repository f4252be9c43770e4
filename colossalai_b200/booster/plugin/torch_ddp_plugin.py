"""TorchDDPPlugin: `torch.nn.parallel.DistributedDataParallel` behind the Booster API.
Parity: reference `colossalai/booster/plugin/torch_ddp_plugin.py:25-340` (checkpoint IO that only writes on rank 0,
`TorchDDPModel`, fp8 gradient compression hook, no_sync, LoRA)."""
from __future__ import annotations

from pathlib import Path
from typing import Callable, Dict, Iterator, List, Optional, Tuple, Union

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.nn.parallel import DistributedDataParallel as DDP
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler as LRScheduler
from torch.utils.data import DataLoader

from ...checkpoint_io import CheckpointIO, GeneralCheckpointIO
from ...cluster import DistCoordinator
from ...interface import ModelWrapper, OptimizerWrapper
from ...accelerator import get_accelerator
from .plugin_base import DPPluginBase

__all__ = ["TorchDDPPlugin", "TorchDDPModel", "TorchDDPCheckpointIO"]


class TorchDDPCheckpointIO(GeneralCheckpointIO):
    """Replicated state: every rank loads, only the master writes."""

    def __init__(self) -> None:
        super().__init__()
        self.coordinator = DistCoordinator()

    def _master(self) -> bool:
        return self.coordinator.is_master()

    def load_unsharded_model(self, model, checkpoint: str, strict: bool = True, low_cpu_mem_mode: bool = True,
                             num_threads: int = 1):
        assert isinstance(model, ModelWrapper), "Please boost the model before loading!"
        super().load_unsharded_model(model.unwrap(), checkpoint, strict, low_cpu_mem_mode, num_threads)

    def save_unsharded_model(self, model, checkpoint: str, gather_dtensor: bool, use_safetensors: bool,
                             use_async: bool = False):
        assert isinstance(model, ModelWrapper), "Please boost the model before saving!"
        if self._master():
            super().save_unsharded_model(model.unwrap(), checkpoint, gather_dtensor, use_safetensors, use_async)

    def load_unsharded_optimizer(self, optimizer, checkpoint: str, low_cpu_mem_mode: bool = True, num_threads: int = 1):
        assert isinstance(optimizer, OptimizerWrapper), "Please boost the optimizer before loading!"
        super().load_unsharded_optimizer(optimizer, checkpoint, low_cpu_mem_mode, num_threads)

    def save_unsharded_optimizer(self, optimizer, checkpoint: str, gather_dtensor: bool, use_async: bool = False):
        assert isinstance(optimizer, OptimizerWrapper), "Please boost the optimizer before saving!"
        if self._master():
            super().save_unsharded_optimizer(optimizer, checkpoint, gather_dtensor, use_async)

    def save_lr_scheduler(self, lr_scheduler: LRScheduler, checkpoint: str):
        if self._master():
            super().save_lr_scheduler(lr_scheduler, checkpoint)

    def save_sharded_model(self, model, checkpoint_path: str, gather_dtensor: bool = True, prefix: Optional[str] = None,
                           max_shard_size: int = 1024, use_safetensors: bool = False, use_async: bool = False):
        assert isinstance(model, ModelWrapper), "Please boost the model before saving!"
        if self._master():
            super().save_sharded_model(model.unwrap(), checkpoint_path, gather_dtensor, prefix, max_shard_size,
                                       use_safetensors, use_async)

    def load_sharded_model(self, model, checkpoint_index_file: str, strict: bool = False, use_safetensors: bool = False,
                           load_sub_module: bool = True, low_cpu_mem_mode: bool = True, num_threads: int = 1):
        assert isinstance(model, ModelWrapper), "Please boost the model before loading!"
        super().load_sharded_model(model.unwrap(), checkpoint_index_file, strict, use_safetensors, load_sub_module,
                                   low_cpu_mem_mode, num_threads)

    def save_sharded_optimizer(self, optimizer, checkpoint: str, gather_dtensor: bool = True, prefix: Optional[str] = None,
                               size_per_shard: int = 1024, use_async: bool = False):
        assert isinstance(optimizer, OptimizerWrapper), "Please boost the optimizer before saving!"
        if self._master():
            super().save_sharded_optimizer(optimizer.unwrap(), checkpoint, gather_dtensor, prefix, size_per_shard,
                                           use_async)

    def load_sharded_optimizer(self, optimizer, index_file_path: str, prefix: Optional[str] = None,
                               low_cpu_mem_mode: bool = True, num_threads: int = 1):
        assert isinstance(optimizer, OptimizerWrapper), "Please boost the optimizer before loading!"
        super().load_sharded_optimizer(optimizer.unwrap(), index_file_path, prefix, low_cpu_mem_mode, num_threads)

    def save_lora_as_pretrained(self, model, checkpoint: str, use_safetensors: bool = False,
                                state_dict: Optional[dict] = None) -> None:
        if self._master():
            super().save_lora_as_pretrained(model, checkpoint, use_safetensors, state_dict)


class TorchDDPModel(ModelWrapper):
    def __init__(self, module: nn.Module, *args, **kwargs) -> None:
        super().__init__(module)
        self.module = DDP(module, *args, **kwargs)

    def unwrap(self, unwrap_peft: bool = True) -> nn.Module:
        return self.module.module


class TorchDDPPlugin(DPPluginBase):
    """
    ```python
    plugin = TorchDDPPlugin()
    booster = Booster(plugin=plugin)
    model, optimizer, criterion, dataloader, _ = booster.boost(model, optimizer, criterion, dataloader)
    ```
    """

    def __init__(self, broadcast_buffers: bool = True, bucket_cap_mb: int = 25, find_unused_parameters: bool = False,
                 check_reduction: bool = False, gradient_as_bucket_view: bool = False, static_graph: bool = False,
                 fp8_communication: bool = False) -> None:
        super().__init__()
        self.ddp_kwargs = dict(broadcast_buffers=broadcast_buffers, bucket_cap_mb=bucket_cap_mb,
                               find_unused_parameters=find_unused_parameters,
                               gradient_as_bucket_view=gradient_as_bucket_view, static_graph=static_graph)
        self.fp8_communication = fp8_communication

    def support_no_sync(self) -> bool:
        return True

    def support_lora(self) -> bool:
        return True

    def control_precision(self) -> bool:
        return False

    def supported_precisions(self) -> List[str]:
        return ["fp16", "fp16_apex", "bf16", "fp8"]

    def control_device(self) -> bool:
        return True

    def supported_devices(self) -> List[str]:
        return ["cuda", "cpu"]

    def configure(self, model: nn.Module, optimizer: Optional[Optimizer] = None, criterion: Optional[Callable] = None,
                  dataloader: Optional[DataLoader] = None, lr_scheduler: Optional[LRScheduler] = None
                  ) -> Tuple[nn.Module, OptimizerWrapper, Callable, DataLoader, LRScheduler]:
        dev = get_accelerator().get_current_device()
        model = model.to(dev)
        # one process may hold exactly one replica -> sync BN works out of the box
        model = nn.SyncBatchNorm.convert_sync_batchnorm(model, None)
        kw = dict(self.ddp_kwargs)
        if dev.type == "cuda":
            kw["device_ids"] = [dev.index if dev.index is not None else torch.cuda.current_device()]
        model = TorchDDPModel(model, **kw)
        if optimizer is not None and not isinstance(optimizer, OptimizerWrapper):
            optimizer = OptimizerWrapper(optimizer)
        if self.fp8_communication:
            from ...quantization.fp8 import fp8_compress_ddp_grad_comm_hook_async

            model.module.register_comm_hook(None, fp8_compress_ddp_grad_comm_hook_async)
        return model, optimizer, criterion, dataloader, lr_scheduler

    def control_checkpoint_io(self) -> bool:
        return True

    def get_checkpoint_io(self) -> CheckpointIO:
        return TorchDDPCheckpointIO()

    def no_sync(self, model: nn.Module, optimizer: OptimizerWrapper) -> Iterator[None]:
        assert isinstance(model, TorchDDPModel), "Model is not boosted by TorchDDPPlugin."
        return model.module.no_sync()

    def enable_lora(self, model: nn.Module, pretrained_dir: Optional[str] = None, lora_config: Optional[Dict] = None,
                    bnb_quantization_config=None) -> nn.Module:
        from ..lora import apply_lora

        if bnb_quantization_config is not None:
            from ...quantization import quantize_model

            model = quantize_model(model, bnb_quantization_config)
        return apply_lora(model, lora_config, pretrained_dir)
