"""GeminiPlugin: chunked ZeRO-3 with heterogeneous memory placement (optionally combined with tensor parallelism).
Parity: reference `colossalai/booster/plugin/gemini_plugin.py:369-712` (+ `GeminiCheckpointIO`)."""
from __future__ import annotations

import gc
import os
from pathlib import Path
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler as LRScheduler
from torch.utils.data import DataLoader

from ...accelerator import get_accelerator
from ...checkpoint_io import CheckpointIndexFile, CheckpointIO, GeneralCheckpointIO
from ...checkpoint_io.utils import (
    get_model_base_filenames,
    get_optimizer_base_filenames,
    load_shard_state_dict,
    load_state_dict,
    save_param_groups,
    save_state_dict,
    save_state_dict_shards,
)
from ...cluster import DeviceMesh, DistCoordinator
from ...interface import ModelWrapper, OptimizerWrapper
from ...logging import get_dist_logger
from ...shardformer import ShardConfig, ShardFormer
from ...zero.gemini import GeminiDDP, GeminiOptimizer
from ...zero.gemini.memory_tracer import MemStats
from .plugin_base import DPPluginBase, _seed_worker

__all__ = ["GeminiPlugin", "GeminiCheckpointIO"]

SUPPORTED_PRECISION = ["fp16", "bf16"]
PRECISION_STR_TO_DTYPE = {"fp16": torch.half, "bf16": torch.bfloat16}


class GeminiCheckpointIO(GeneralCheckpointIO):
    def __init__(self) -> None:
        super().__init__()
        self.coordinator = DistCoordinator()

    def save_unsharded_model(self, model: GeminiDDP, checkpoint: str, gather_dtensor: bool, use_safetensors: bool,
                             use_async: bool = False):
        assert isinstance(model, GeminiDDP), "Please boost the model before saving!"
        sd = model.state_dict(only_rank_0=True)
        if self.coordinator.is_master():
            save_state_dict(sd, checkpoint, use_safetensors)
        dist.barrier()

    def load_unsharded_model(self, model: GeminiDDP, checkpoint: str, strict: bool = True,
                             low_cpu_mem_mode: bool = True, num_threads: int = 1):
        assert isinstance(model, GeminiDDP), "Please boost the model before loading!"
        model.load_state_dict(load_state_dict(checkpoint), strict=strict)

    def save_sharded_model(self, model: GeminiDDP, checkpoint_path: str, gather_dtensor: bool = False,
                           prefix: Optional[str] = None, max_shard_size: int = 1024, use_safetensors: bool = False,
                           use_async: bool = False):
        assert isinstance(model, GeminiDDP), "Please boost the model before saving!"
        if os.path.isfile(checkpoint_path):
            return
        Path(checkpoint_path).mkdir(parents=True, exist_ok=True)
        shards = model.state_dict_shard(max_shard_size=max_shard_size, only_rank_0=True, dtype=model.mixed_precision)
        weights_name, save_index_file = get_model_base_filenames(prefix, use_safetensors)
        index_file = CheckpointIndexFile(checkpoint_path)
        is_master = self.coordinator.is_master()
        total = save_state_dict_shards(shards, checkpoint_path, index_file, weights_name, is_master, use_safetensors)
        if is_master:
            index_file.append_meta_data("total_size", total)
            index_file.write_index_file(save_index_file)
        dist.barrier()

    def load_sharded_model(self, model: GeminiDDP, checkpoint_index_file: Path, strict: bool = False,
                           use_safetensors: bool = False, load_sub_module: bool = True, low_cpu_mem_mode: bool = True,
                           num_threads: int = 1):
        assert isinstance(model, GeminiDDP), "Please boost the model before loading!"
        idx = CheckpointIndexFile.from_file(checkpoint_index_file)
        sd = {}
        for fn in idx.get_checkpoint_filenames():
            sd.update(load_shard_state_dict(Path(fn)))
        model.load_state_dict(sd, strict=strict)

    def save_unsharded_optimizer(self, optimizer: GeminiOptimizer, checkpoint: str, gather_dtensor: bool,
                                 use_async: bool = False):
        assert isinstance(optimizer, GeminiOptimizer), "Please boost the optimizer before saving!"
        sd = optimizer.state_dict()
        if self.coordinator.is_master():
            torch.save(sd, checkpoint)
        dist.barrier()

    def load_unsharded_optimizer(self, optimizer: GeminiOptimizer, checkpoint: str, low_cpu_mem_mode: bool = True,
                                 num_threads: int = 1):
        assert isinstance(optimizer, GeminiOptimizer), "Please boost the optimizer before loading!"
        optimizer.load_state_dict(load_state_dict(checkpoint))

    def save_sharded_optimizer(self, optimizer: GeminiOptimizer, checkpoint: Path, gather_dtensor: bool, prefix: str,
                               size_per_shard: int, use_async: bool = False):
        assert isinstance(optimizer, GeminiOptimizer), "Please boost the optimizer before saving!"
        Path(checkpoint).mkdir(parents=True, exist_ok=True)
        sd = optimizer.state_dict()
        if self.coordinator.is_master():
            states_name, save_index_file, param_group_file = get_optimizer_base_filenames(prefix)
            index_file = CheckpointIndexFile(checkpoint)
            index_file.append_meta_data("param_groups", param_group_file)
            save_param_groups(sd, os.path.join(checkpoint, param_group_file))
            from ...checkpoint_io.utils import shard_optimizer_checkpoint

            total = save_state_dict_shards(shard_optimizer_checkpoint(sd, size_per_shard), checkpoint, index_file,
                                           states_name, True)
            index_file.append_meta_data("total_size", total)
            index_file.write_index_file(save_index_file)
        dist.barrier()

    def load_sharded_optimizer(self, optimizer: GeminiOptimizer, checkpoint_index_file: Path, prefix: str,
                               low_cpu_mem_mode: bool = True, num_threads: int = 1):
        assert isinstance(optimizer, GeminiOptimizer), "Please boost the optimizer before loading!"
        idx = CheckpointIndexFile.from_file(checkpoint_index_file)
        groups = torch.load(idx.get_param_group_filename(), weights_only=False)
        state = {}
        for fn in idx.get_checkpoint_filenames():
            state.update(load_shard_state_dict(Path(fn)))
        optimizer.load_state_dict({"state": state, "param_groups": groups})

    def save_lr_scheduler(self, lr_scheduler: LRScheduler, checkpoint: str):
        if self.coordinator.is_master():
            super().save_lr_scheduler(lr_scheduler, checkpoint)


class GeminiPlugin(DPPluginBase):
    """
    >>> plugin = GeminiPlugin(placement_policy="static", shard_param_frac=1.0, offload_optim_frac=1.0,
    ...                       precision="bf16", pin_memory=True, max_norm=1.0)
    >>> model, optimizer, *_ = Booster(plugin=plugin).boost(model, HybridAdam(model.parameters()))
    """

    def __init__(self, chunk_config_dict: Optional[dict] = None,
                 chunk_init_device: Optional[torch.device] = None, placement_policy: str = "static",
                 enable_gradient_accumulation: bool = False, max_prefetch: int = 0, shard_param_frac: float = 1.0,
                 offload_optim_frac: float = 0.0, offload_param_frac: float = 0.0,
                 warmup_non_model_data_ratio: float = 0.8, steady_cuda_cap_ratio: float = 0.9,
                 precision: str = "fp16", master_weights: bool = True, pin_memory: bool = False,
                 force_outputs_fp32: bool = False, strict_ddp_mode: bool = False, search_range_m: int = 32,
                 hidden_dim: Optional[int] = None, min_chunk_size_m: float = 32, memstats: Optional[MemStats] = None,
                 gpu_margin_mem_ratio: float = 0.0, initial_scale: float = 2**16, min_scale: float = 1,
                 growth_factor: float = 2, backoff_factor: float = 0.5, growth_interval: int = 1000,
                 hysteresis: int = 2, max_scale: float = 2**32, max_norm: float = 0.0, norm_type: float = 2.0,
                 tp_size: int = 1, extra_dp_size: int = 1, enable_all_optimization: bool = False,
                 enable_fused_normalization: bool = False, enable_flash_attention: bool = False,
                 enable_sequence_parallelism: bool = False, enable_jit_fused: bool = False,
                 enable_async_reduce: bool = True, use_fp8: bool = False, verbose: bool = False,
                 fp8_communication: bool = False) -> None:
        super().__init__()
        assert precision in SUPPORTED_PRECISION, f"precision {precision} is not supported"
        self.gemini_config = dict(
            chunk_config_dict=chunk_config_dict,
            chunk_init_device=chunk_init_device or get_accelerator().get_current_device(),
            placement_policy=placement_policy, enable_gradient_accumulation=enable_gradient_accumulation,
            shard_param_frac=shard_param_frac, offload_optim_frac=offload_optim_frac,
            offload_param_frac=offload_param_frac, warmup_non_model_data_ratio=warmup_non_model_data_ratio,
            steady_cuda_cap_ratio=steady_cuda_cap_ratio, pin_memory=pin_memory, force_outputs_fp32=force_outputs_fp32,
            strict_ddp_mode=strict_ddp_mode, search_range_m=search_range_m, hidden_dim=hidden_dim,
            min_chunk_size_m=min_chunk_size_m, memstats=memstats, mixed_precision=PRECISION_STR_TO_DTYPE[precision],
            master_weights=master_weights, max_prefetch=max_prefetch, enable_async_reduce=enable_async_reduce,
            fp8_communication=fp8_communication, use_fp8=use_fp8)
        self.zero_optim_config = dict(gpu_margin_mem_ratio=gpu_margin_mem_ratio)
        self.optim_kwargs = dict(initial_scale=initial_scale, growth_factor=growth_factor,
                                 backoff_factor=backoff_factor, growth_interval=growth_interval, hysteresis=hysteresis,
                                 min_scale=min_scale, max_scale=max_scale, max_norm=max_norm, norm_type=norm_type)
        self.enable_tensor_parallelism = tp_size > 1
        self.enable_all_optimization = enable_all_optimization
        self.enable_fused_normalization = enable_fused_normalization
        self.enable_flash_attention = enable_flash_attention
        self.enable_sequence_parallelism = enable_sequence_parallelism if self.enable_tensor_parallelism else False
        self.enable_jit_fused = enable_jit_fused
        self.verbose = verbose
        self.tp_size, self.extra_dp_size = tp_size, extra_dp_size
        world_size = dist.get_world_size()
        self.zero_size = world_size // (self.tp_size * self.extra_dp_size)
        assert world_size == self.tp_size * self.extra_dp_size * self.zero_size, (
            f"The global group size can't be evenly divided by the subgroup size.")
        self.pg_mesh = DeviceMesh(zero=self.zero_size, extra_dp=self.extra_dp_size, tp=self.tp_size)
        self.zero_group = self.pg_mesh.group("zero") if self.zero_size < world_size else None
        self.extra_dp_group = self.pg_mesh.group("extra_dp") if self.extra_dp_size > 1 else None
        self.tp_group = self.pg_mesh.group("tp") if self.tp_size > 1 else None
        self.dp_size = self.zero_size * self.extra_dp_size
        self.shard_config = ShardConfig(
            tensor_parallel_process_group=self.tp_group, enable_tensor_parallelism=self.enable_tensor_parallelism,
            enable_all_optimization=self.enable_all_optimization,
            enable_fused_normalization=self.enable_fused_normalization,
            enable_flash_attention=self.enable_flash_attention, enable_jit_fused=self.enable_jit_fused,
            enable_sequence_parallelism=self.enable_sequence_parallelism)
        self.logger = get_dist_logger()

    def __del__(self):
        try:
            self.pg_mesh.destroy_mesh_process_groups()
        except Exception:
            pass

    def support_no_sync(self) -> bool:
        return False

    def support_lora(self) -> bool:
        return False

    def control_precision(self) -> bool:
        return True

    def supported_precisions(self) -> List[str]:
        return SUPPORTED_PRECISION

    def control_device(self) -> bool:
        return True

    def supported_devices(self) -> List[str]:
        return ["cuda", "cpu"]

    def prepare_dataloader(self, dataset, batch_size, shuffle=False, seed=1024, drop_last=False, pin_memory=False,
                           num_workers=0, distributed_sampler_cls=None, **kwargs):
        from torch.utils.data.distributed import DistributedSampler

        cls = distributed_sampler_cls or DistributedSampler
        zero_rank, extra_rank = self.pg_mesh.axis_rank("zero"), self.pg_mesh.axis_rank("extra_dp")
        sampler = cls(dataset, num_replicas=self.dp_size, rank=extra_rank * self.zero_size + zero_rank, shuffle=shuffle)
        return DataLoader(dataset, batch_size=batch_size, sampler=sampler, worker_init_fn=_seed_worker(seed),
                          drop_last=drop_last, pin_memory=pin_memory, num_workers=num_workers, **kwargs)

    def configure(self, model: nn.Module, optimizer: Optional[Optimizer] = None, criterion: Optional[Callable] = None,
                  dataloader: Optional[DataLoader] = None, lr_scheduler: Optional[LRScheduler] = None):
        params_info = None
        if not isinstance(model, ModelWrapper):
            if self.enable_tensor_parallelism:
                model, _ = ShardFormer(self.shard_config).optimize(model)
            orig_params = list(model.parameters())
            model = GeminiDDP(model, **self.gemini_config, zero_group=self.zero_group,
                              extra_dp_group=self.extra_dp_group, verbose=self.verbose)
            if optimizer is not None and not isinstance(optimizer, OptimizerWrapper):
                live = [p for p in model.fp16_params if p.requires_grad]
                if len(optimizer.param_groups) == 1:
                    optimizer.param_groups[0]["params"] = live
                else:
                    live_ids = {id(p) for p in live}
                    for g in optimizer.param_groups:
                        g["params"] = [p for p in g["params"] if id(p) in live_ids]
        if optimizer is not None and not isinstance(optimizer, OptimizerWrapper):
            optimizer = GeminiOptimizer(optimizer, model, **self.zero_optim_config, **self.optim_kwargs,
                                        tp_group=self.tp_group, params_info=params_info, verbose=self.verbose)
        return model, optimizer, criterion, dataloader, lr_scheduler

    def control_checkpoint_io(self) -> bool:
        return True

    def get_checkpoint_io(self) -> CheckpointIO:
        return GeminiCheckpointIO()

    def no_sync(self, model: nn.Module, optimizer: OptimizerWrapper) -> Iterator[None]:
        raise NotImplementedError("GeminiPlugin does not support no_sync; use enable_gradient_accumulation")

    def enable_lora(self, model, pretrained_dir=None, lora_config=None, bnb_quantization_config=None):
        raise NotImplementedError("GeminiPlugin does not support LoRA")
