"""Engine ABC.  Parity: reference `colossalai/inference/core/base_engine.py`."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
import torch.nn as nn

from ...cluster import ProcessGroupMesh

__all__ = ["BaseEngine"]


class BaseEngine(ABC):
    @abstractmethod
    def __init__(self, model_or_path, inference_config=None, verbose: bool = False, model_policy=None):
        ...

    @abstractmethod
    def init_model(self, model_or_path, model_policy=None, model_shard_infer_config=None):
        ...

    @abstractmethod
    def generate(self, request_ids=None, prompts=None, generation_config=None, **kwargs):
        ...

    @abstractmethod
    def add_request(self, prompts, request_ids=None, **kwargs):
        ...

    @abstractmethod
    def step(self):
        ...

    @abstractmethod
    def _verify_args(self):
        ...

    @torch.inference_mode()
    def capture_model(self):
        return NotImplementedError("This method should be implemented by subclasses")

    def _shardformer(self, model: nn.Module, model_policy, model_shard_infer_config=None, stage_manager=None,
                     tp_group=None, **kwargs) -> nn.Module:
        """Shard `model` over the TP group with our policy machinery."""
        from ...shardformer import ShardConfig, ShardFormer

        sc = ShardConfig(tensor_parallel_process_group=tp_group, pipeline_stage_manager=stage_manager,
                         enable_tensor_parallelism=(tp_group is not None), enable_fused_normalization=False,
                         enable_flash_attention=False, enable_jit_fused=False, enable_sequence_parallelism=False,
                         parallel_output=False, extra_kwargs={"model_shard_infer_config": model_shard_infer_config, **kwargs})
        model, _ = ShardFormer(sc).optimize(model, model_policy)
        return model
