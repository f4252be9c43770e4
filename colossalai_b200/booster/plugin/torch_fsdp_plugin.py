"""TorchFSDPPlugin: `torch.distributed.fsdp.FullyShardedDataParallel` behind the Booster API.
Parity: reference `colossalai/booster/plugin/torch_fsdp_plugin.py:40-580` (full-state-dict checkpoint IO on rank 0,
sharded save via the full state, `TorchFSDPModel`, `FSDPOptimizerWrapper`, fp8 comm hooks)."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler as LRScheduler
from torch.utils.data import DataLoader

from ...accelerator import get_accelerator
from ...checkpoint_io import CheckpointIndexFile, CheckpointIO, GeneralCheckpointIO
from ...checkpoint_io import utils as ckpt_utils
from ...cluster import DistCoordinator
from ...interface import ModelWrapper, OptimizerWrapper
from .plugin_base import DPPluginBase

__all__ = ["TorchFSDPPlugin", "TorchFSDPModel", "TorchFSDPCheckpointIO", "FSDPOptimizerWrapper"]


def _fsdp():
    from torch.distributed.fsdp import FullStateDictConfig, FullyShardedDataParallel as FSDP, StateDictType

    return FSDP, StateDictType, FullStateDictConfig


class TorchFSDPCheckpointIO(GeneralCheckpointIO):
    def __init__(self) -> None:
        super().__init__()
        self.coordinator = DistCoordinator()

    # ---- model
    def _full_model_state(self, model) -> dict:
        FSDP, SDT, Cfg = _fsdp()
        with FSDP.state_dict_type(model.unwrap(), SDT.FULL_STATE_DICT, Cfg(offload_to_cpu=True, rank0_only=True)):
            return model.unwrap().state_dict()

    def load_unsharded_model(self, model, checkpoint: str, strict: bool = True, low_cpu_mem_mode: bool = True,
                             num_threads: int = 1):
        assert isinstance(model, TorchFSDPModel), "Please boost the model before loading!"
        FSDP, SDT, Cfg = _fsdp()
        sd = ckpt_utils.load_state_dict(checkpoint)
        with FSDP.state_dict_type(model.unwrap(), SDT.FULL_STATE_DICT, Cfg(offload_to_cpu=True, rank0_only=False)):
            model.unwrap().load_state_dict(sd, strict=strict)

    def save_unsharded_model(self, model, checkpoint: str, gather_dtensor: bool, use_safetensors: bool,
                             use_async: bool = False):
        assert isinstance(model, TorchFSDPModel), "Please boost the model before saving!"
        sd = self._full_model_state(model)
        if self.coordinator.is_master():
            ckpt_utils.save_state_dict(sd, checkpoint, use_safetensors)

    def save_sharded_model(self, model, checkpoint_path: str, gather_dtensor: bool = True, prefix: Optional[str] = None,
                           size_per_shard: int = 1024, use_safetensors: bool = False, use_async: bool = False):
        assert isinstance(model, TorchFSDPModel), "Please boost the model before saving!"
        if os.path.isfile(checkpoint_path):
            raise ValueError(f"Provided path ({checkpoint_path}) should be a directory, not a file")
        Path(checkpoint_path).mkdir(parents=True, exist_ok=True)
        sd = self._full_model_state(model)
        if not self.coordinator.is_master():
            return
        weights_name, index_name = ckpt_utils.get_model_base_filenames(prefix, use_safetensors)
        index = CheckpointIndexFile(checkpoint_path)
        shards = ckpt_utils.shard_model_checkpoint(sd, max_shard_size=size_per_shard)
        total = ckpt_utils.save_state_dict_shards(shards, checkpoint_path, index, weights_name, True, use_safetensors)
        index.append_meta_data("total_size", total)
        index.write_index_file(index_name)

    def load_sharded_model(self, model, checkpoint_index_file: Path, strict: bool = False, use_safetensors: bool = False,
                           load_sub_module: bool = True, low_cpu_mem_mode: bool = True, num_threads: int = 1):
        assert isinstance(model, TorchFSDPModel), "Please boost the model before loading!"
        FSDP, SDT, Cfg = _fsdp()
        index = CheckpointIndexFile.from_file(checkpoint_index_file)
        sd = {}
        for f in index.get_checkpoint_filenames():
            sd.update(ckpt_utils.load_shard_state_dict(Path(f), ckpt_utils.is_safetensor_checkpoint(f)))
        with FSDP.state_dict_type(model.unwrap(), SDT.FULL_STATE_DICT, Cfg(offload_to_cpu=True, rank0_only=False)):
            model.unwrap().load_state_dict(sd, strict=False)

    # ---- optimizer
    def save_unsharded_optimizer(self, optimizer, checkpoint: str, gather_dtensor: bool, use_async: bool = False):
        assert isinstance(optimizer, FSDPOptimizerWrapper), "Please boost the optimizer before saving!"
        FSDP, *_ = _fsdp()
        sd = FSDP.full_optim_state_dict(optimizer.unwrap_model().unwrap(), optim=optimizer.unwrap(), rank0_only=True)
        if self.coordinator.is_master():
            torch.save(sd, checkpoint)

    def load_unsharded_optimizer(self, optimizer, checkpoint: str, low_cpu_mem_mode: bool = True, num_threads: int = 1):
        assert isinstance(optimizer, FSDPOptimizerWrapper), "Please boost the optimizer before loading!"
        FSDP, *_ = _fsdp()
        full = torch.load(checkpoint, map_location="cpu", weights_only=False)
        fsdp_model = optimizer.unwrap_model().unwrap()
        sharded = FSDP.scatter_full_optim_state_dict(full, fsdp_model, optim=optimizer.unwrap())
        optimizer.unwrap().load_state_dict(sharded)

    def save_sharded_optimizer(self, optimizer, checkpoint: str, gather_dtensor: bool = True, prefix: Optional[str] = None,
                               size_per_shard: int = 1024, use_async: bool = False):
        Path(checkpoint).mkdir(parents=True, exist_ok=True)
        self.save_unsharded_optimizer(optimizer, os.path.join(checkpoint, (prefix or "") + "optimizer.bin"),
                                      gather_dtensor)

    def load_sharded_optimizer(self, optimizer, index_file_path: str, prefix: Optional[str] = None,
                               low_cpu_mem_mode: bool = True, num_threads: int = 1):
        d = index_file_path if os.path.isdir(index_file_path) else os.path.dirname(index_file_path)
        self.load_unsharded_optimizer(optimizer, os.path.join(d, (prefix or "") + "optimizer.bin"))

    def save_lr_scheduler(self, lr_scheduler: LRScheduler, checkpoint: str):
        if self.coordinator.is_master():
            super().save_lr_scheduler(lr_scheduler, checkpoint)


class TorchFSDPModel(ModelWrapper):
    def __init__(self, module: nn.Module, *args, **kwargs) -> None:
        super().__init__(module)
        FSDP, *_ = _fsdp()
        self.module = FSDP(module, *args, **kwargs)


class FSDPOptimizerWrapper(OptimizerWrapper):
    def __init__(self, optimizer: Optimizer, model: nn.Module) -> None:
        self.model = model
        super().__init__(optimizer)

    def unwrap_model(self) -> nn.Module:
        return self.model


class TorchFSDPPlugin(DPPluginBase):
    """
    ```python
    plugin = TorchFSDPPlugin()
    booster = Booster(plugin=plugin)
    model, optimizer, *_ = booster.boost(model, optimizer)
    ```
    """

    def __init__(self, process_group=None, sharding_strategy=None, cpu_offload=None, auto_wrap_policy=None,
                 backward_prefetch=None, mixed_precision=None, ignored_modules: Optional[Iterable[nn.Module]] = None,
                 param_init_fn: Optional[Callable[[nn.Module], None]] = None, sync_module_states: bool = False,
                 fp8_communication: bool = False) -> None:
        super().__init__()
        self.fsdp_kwargs = dict(process_group=process_group, sharding_strategy=sharding_strategy,
                                cpu_offload=cpu_offload, auto_wrap_policy=auto_wrap_policy,
                                backward_prefetch=backward_prefetch, mixed_precision=mixed_precision,
                                ignored_modules=ignored_modules, param_init_fn=param_init_fn,
                                sync_module_states=sync_module_states)
        self.fp8_communication = fp8_communication

    def support_no_sync(self) -> bool:
        return False

    def support_lora(self) -> bool:
        return False

    def no_sync(self, model: nn.Module, optimizer: OptimizerWrapper) -> Iterator[None]:
        raise NotImplementedError("Torch fsdp no_sync func not supported yet.")

    def control_precision(self) -> bool:
        return True

    def supported_precisions(self) -> List[str]:
        return ["fp16", "bf16"]

    def control_device(self) -> bool:
        return True

    def supported_devices(self) -> List[str]:
        return ["cuda", "cpu"]

    def configure(self, model: nn.Module, optimizer: Optional[Optimizer] = None, criterion: Optional[Callable] = None,
                  dataloader: Optional[DataLoader] = None, lr_scheduler: Optional[LRScheduler] = None
                  ) -> Tuple[nn.Module, OptimizerWrapper, Callable, DataLoader, LRScheduler]:
        dev = get_accelerator().get_current_device()
        kw = {k: v for k, v in self.fsdp_kwargs.items() if v is not None and v is not False}
        if dev.type == "cuda":
            kw["device_id"] = torch.cuda.current_device()
        fsdp_model = TorchFSDPModel(model, **kw)
        if self.fp8_communication:
            from ...quantization.fp8 import fp8_compress_fsdp_grad_comm_hook, fp8_compress_fsdp_params_comm_hook
            from ...quantization.utils import patch_fsdp_params_comm_hook

            patch_fsdp_params_comm_hook()
            fsdp_model.module.register_params_comm_hook(None, fp8_compress_fsdp_params_comm_hook)
            fsdp_model.module.register_comm_hook(None, fp8_compress_fsdp_grad_comm_hook)
        if optimizer is not None:
            if len(optimizer.param_groups) > 1:
                import warnings

                warnings.warn("TorchFSDPPlugin does not support optimizers that use multi param groups; the optimizer "
                              "is re-initialised over the flattened FSDP parameters.")
            # the wrapped module owns NEW flat parameters -> rebuild the optimizer over them, keep the hyper-parameters
            defaults = {k: v for k, v in optimizer.defaults.items()}
            optimizer.__init__(fsdp_model.parameters(), **defaults)
            if not isinstance(optimizer, FSDPOptimizerWrapper):
                optimizer = FSDPOptimizerWrapper(optimizer, fsdp_model)
        return fsdp_model, optimizer, criterion, dataloader, lr_scheduler

    def control_checkpoint_io(self) -> bool:
        return True

    def get_checkpoint_io(self) -> CheckpointIO:
        return TorchFSDPCheckpointIO()

    def enable_lora(self, model: nn.Module, pretrained_dir: Optional[str] = None, lora_config: Optional[Dict] = None,
                    bnb_quantization_config=None) -> nn.Module:
        raise NotImplementedError("TorchFSDPPlugin does not support LoRA")
