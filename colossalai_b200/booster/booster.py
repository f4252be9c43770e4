"""Booster façade.  Parity: reference `colossalai/booster/booster.py:33-433`."""
from __future__ import annotations

import os

from contextlib import contextmanager
from typing import Any, Callable, Dict, Iterator, List, Optional, Union

import torch
import torch.nn as nn
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler as LRScheduler
from torch.utils.data import DataLoader

from ..checkpoint_io import GeneralCheckpointIO
from ..interface import ModelWrapper, OptimizerWrapper
from ..logging import get_dist_logger
from .accelerator import Accelerator
from .mixed_precision import MixedPrecision, mixed_precision_factory
from .plugin.plugin_base import Plugin

__all__ = ["Booster"]


class Booster:
    """
    >>> colossalai_b200.launch_from_torch()
    >>> booster = Booster(plugin=HybridParallelPlugin(tp_size=8, pp_size=1, precision="bf16"))
    >>> model, optimizer, criterion, dataloader, lr_scheduler = booster.boost(model, optimizer, criterion, dataloader)
    >>> for batch in dataloader:
    ...     loss = model(**batch)["loss"]
    ...     booster.backward(loss, optimizer)
    ...     optimizer.step(); optimizer.zero_grad()
    """

    def __init__(self, device: Optional[str] = None, mixed_precision: Optional[Union[MixedPrecision, str]] = None,
                 plugin: Optional[Plugin] = None, convert_hf_models: bool = True) -> None:
        """`convert_hf_models`: a `transformers` model handed to `boost` is converted weight-for-weight into the native
        model of its family (every parallelism, fused kernels).  With False the user's module is kept and sharded IN
        PLACE by the HuggingFace policies of `shardformer/policies/hf_*.py` (tensor / expert parallelism, ZeRO, DDP;
        no pipeline or sequence parallelism), like the reference does."""
        self.convert_hf_models = convert_hf_models
        if plugin is not None:
            assert isinstance(plugin, Plugin), f"plugin must be a Plugin, got {type(plugin)}"
        self.plugin = plugin
        self.logger = get_dist_logger()
        if self.plugin and self.plugin.control_device():
            self.accelerator = None
            if device is not None:
                self.logger.warning("The plugin will control the accelerator, so the device argument will be ignored.",
                                    ranks=[0])
        else:
            if device is None:
                device = "cuda" if torch.cuda.is_available() else "cpu"
            self.accelerator = Accelerator(device)
        if self.plugin and self.plugin.control_precision():
            if mixed_precision is not None:
                self.logger.warning("The plugin will control the precision, so the mixed_precision argument will be ignored.",
                                    ranks=[0])
            self.mixed_precision = None
        elif mixed_precision is None:
            self.mixed_precision = None
        else:
            if isinstance(mixed_precision, str):
                self.mixed_precision = mixed_precision_factory(mixed_precision)
            elif isinstance(mixed_precision, MixedPrecision):
                self.mixed_precision = mixed_precision
            else:
                raise ValueError(f"mixed_precision must be a string or a MixedPrecision, got {type(mixed_precision)}")
        if self.plugin is not None and self.plugin.control_checkpoint_io():
            self.checkpoint_io = self.plugin.get_checkpoint_io()
        else:
            self.checkpoint_io = GeneralCheckpointIO()

    # ------------------------------------------------------------------ boost
    def boost(self, model: nn.Module, optimizer: Optional[Optimizer] = None, criterion: Optional[Callable] = None,
              dataloader: Optional[DataLoader] = None, lr_scheduler: Optional[LRScheduler] = None) -> List[Any]:
        from ..interface.pretrained import get_pretrained_path, set_pretrained_path

        # reference-style user code hands over a Hugging Face model instance: convert it to our implementation
        from ..models.hf_io import from_hf_model, is_hf_model

        if is_hf_model(model) and self.convert_hf_models:
            self.logger.info(f"converting {type(model).__name__} (transformers) to the colossalai_b200 model of the "
                             "same family; rebuild the optimizer over the returned model's parameters", ranks=[0])
            if optimizer is not None:
                hf_ids = {id(p) for p in model.parameters()}
                opt_ids = {id(p) for g in optimizer.param_groups for p in g["params"]}
                assert not (opt_ids & hf_ids), (
                    "the optimizer was built over the Hugging Face model's parameters, which are replaced by the "
                    "conversion: call `colossalai_b200.models.hf_io.from_hf_model(model)` first and build the "
                    "optimizer over the converted model")
            model = from_hf_model(model)
        pretrained_path = get_pretrained_path(model)
        # lazily built models: plugins that shard through ShardFormer materialise AFTER sharding (each rank only allocates
        # its slices); every other plugin gets real tensors before it wraps the module
        if not getattr(self.plugin, "materializes_lazy_models", False) and not isinstance(model, ModelWrapper):
            from ..lazy import LazyInitContext
            from ..lazy.lazy_init import is_lazy

            if any(is_lazy(p) or p.device.type == "meta" for p in model.parameters()):
                LazyInitContext.materialize(model)
        if self.plugin:
            model, optimizer, criterion, dataloader, lr_scheduler = self.plugin.configure(
                model, optimizer, criterion, dataloader, lr_scheduler)
        if self.plugin and not self.plugin.control_device():
            model = self.accelerator.configure_model(model)
        if self.mixed_precision and (self.plugin is None or not self.plugin.control_precision()):
            model, optimizer, criterion = self.mixed_precision.configure(model, optimizer, criterion)
        if self.plugin is None:
            model = self.accelerator.configure_model(model) if self.accelerator else model
            if optimizer is not None and not isinstance(optimizer, OptimizerWrapper):
                optimizer = OptimizerWrapper(optimizer)
        if pretrained_path:
            from ..lazy.pretrained import is_hf_checkpoint_dir, load_pretrained_into

            if is_hf_checkpoint_dir(pretrained_path) and not os.path.isfile(
                    os.path.join(pretrained_path, "cb200_format")):
                load_pretrained_into(model, pretrained_path)      # HF naming -> fused / sharded parameters
            else:
                self.load_model(model, pretrained_path)
            orig = model.unwrap() if isinstance(model, ModelWrapper) else model
            set_pretrained_path(orig, None)
        return model, optimizer, criterion, dataloader, lr_scheduler

    def backward(self, loss: torch.Tensor, optimizer: Optimizer) -> None:
        optimizer.backward(loss)
        # plugins whose DP sync is explicit (hybrid without ZeRO) sync here
        plugin = self.plugin
        model = getattr(optimizer, "model", None)
        if plugin is not None and model is not None and hasattr(model, "sync_dp_grads") \
                and getattr(plugin, "zero_stage", 0) == 0 and getattr(model, "require_grad_sync", True):
            model.sync_dp_grads()

    def execute_pipeline(self, data_iter: Iterator, model: nn.Module, criterion: Callable[[Any, Any], torch.Tensor],
                         optimizer: Optional[Optimizer] = None, return_loss: bool = True,
                         return_outputs: bool = False) -> Dict[str, Any]:
        assert hasattr(self.plugin, "execute_pipeline"), (
            f"The plugin {self.plugin.__class__.__name__} does not support pipeline. Please use HybridParallelPlugin.")
        return self.plugin.execute_pipeline(data_iter, model, criterion, optimizer, return_loss, return_outputs)

    def no_sync(self, model: nn.Module = None, optimizer: OptimizerWrapper = None) -> contextmanager:
        assert self.plugin is not None, "no_sync is only enabled when a plugin is provided and the plugin supports no_sync."
        assert self.plugin.support_no_sync(), "The plugin does not support no_sync."
        return self.plugin.no_sync(model, optimizer)

    def enable_lora(self, model: nn.Module, pretrained_dir: Optional[str] = None, lora_config=None,
                    bnb_quantization_config=None, quantize: bool = False) -> nn.Module:
        assert self.plugin is not None, "Lora can only be enabled when a plugin is provided."
        assert self.plugin.support_lora(), f"The plugin {self.plugin.__class__.__name__} does not support lora."
        if pretrained_dir is None:
            assert lora_config is not None, "Please provide configuration for Lora when pretrained directory path isn't passed in."
        return self.plugin.enable_lora(model, pretrained_dir, lora_config, bnb_quantization_config)

    # ------------------------------------------------------------------ checkpoints
    def load_model(self, model: Union[nn.Module, ModelWrapper], checkpoint: str, strict: bool = True,
                   low_cpu_mem_mode: bool = True, num_threads: int = 1) -> None:
        self.checkpoint_io.load_model(model, checkpoint, strict, low_cpu_mem_mode=low_cpu_mem_mode,
                                      num_threads=num_threads)

    def save_model(self, model: Union[nn.Module, ModelWrapper], checkpoint: str, shard: bool = False,
                   gather_dtensor: bool = True, prefix: Optional[str] = None, size_per_shard: int = 1024,
                   use_safetensors: bool = False, use_async: bool = False) -> None:
        self.checkpoint_io.save_model(model, checkpoint=checkpoint, shard=shard, gather_dtensor=gather_dtensor,
                                      prefix=prefix, size_per_shard=size_per_shard, use_safetensors=use_safetensors,
                                      use_async=use_async)

    def load_optimizer(self, optimizer: Optimizer, checkpoint: str, low_cpu_mem_mode: bool = True,
                       num_threads: int = 1) -> None:
        self.checkpoint_io.load_optimizer(optimizer, checkpoint, low_cpu_mem_mode=low_cpu_mem_mode,
                                          num_threads=num_threads)

    def save_optimizer(self, optimizer: Optimizer, checkpoint: str, shard: bool = False, gather_dtensor: bool = True,
                       prefix: Optional[str] = None, size_per_shard: int = 1024, use_async: bool = False) -> None:
        self.checkpoint_io.save_optimizer(optimizer, checkpoint, shard, gather_dtensor, prefix, size_per_shard,
                                          use_async=use_async)

    def save_lr_scheduler(self, lr_scheduler: LRScheduler, checkpoint: str) -> None:
        self.checkpoint_io.save_lr_scheduler(lr_scheduler, checkpoint)

    def load_lr_scheduler(self, lr_scheduler: LRScheduler, checkpoint: str) -> None:
        self.checkpoint_io.load_lr_scheduler(lr_scheduler, checkpoint)

    def save_lora_as_pretrained(self, model: Union[nn.Module, ModelWrapper], checkpoint: str,
                                use_safetensors: bool = False) -> None:
        if not self.plugin or not self.plugin.support_lora():
            raise ValueError("the current plugin does not support LoRA")
        self.checkpoint_io.save_lora_as_pretrained(model, checkpoint, use_safetensors)
