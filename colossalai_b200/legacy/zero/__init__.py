"""The first-generation ZeRO API (reference `colossalai/legacy/zero`): `ZeroInitContext` (shard parameters at
construction), `ShardedModelV2`, `ShardedOptimizerV2`, shard strategies and `convert_to_zero_v2`.

The old implementation hooked every sub-module to gather / release `ShardedParamV2`s; the chunked Gemini runtime
(`colossalai_b200.zero.gemini`) supersedes it, so the legacy classes are thin adapters that keep the old constructor
arguments and map them onto chunks (`tensor_placement_policy` -> Gemini placement, `reduce_scatter_bucket_size_mb` ->
chunk size, `reuse_fp16_shard` / `gpu_margin_mem_ratio` accepted for compatibility).

Parity: `legacy/zero/init_ctx/init_context.py:1-270`, `legacy/zero/sharded_model/sharded_model_v2.py:1-580`,
`legacy/zero/sharded_optim/sharded_optim_v2.py:1-400`, `legacy/zero/shard_utils/*.py`, `legacy/zero/__init__.py`."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, Optional

import torch
import torch.nn as nn

from ...utils.model.utils import InsertPostInitMethodToModuleSubClasses
from ...zero.gemini import GeminiDDP
from ...zero.gemini.gemini_optimizer import GeminiOptimizer

__all__ = ["TensorShardStrategy", "BucketTensorShardStrategy", "ZeroInitContext", "ShardedModelV2",
           "ShardedOptimizerV2", "convert_to_zero_v2", "no_shard_zero_context", "no_shard_zero_decrator"]


class TensorShardStrategy:
    """Every parameter is flattened and split evenly over the data-parallel ranks (one collective per tensor)."""

    bucketed = False

    def chunk_mb(self, bucket_mb: int) -> int:
        return max(1, min(bucket_mb, 8))


class BucketTensorShardStrategy(TensorShardStrategy):
    """Like `TensorShardStrategy` but gathers many tensors with one collective — the native behaviour of chunks."""

    bucketed = True

    def chunk_mb(self, bucket_mb: int) -> int:
        return max(1, bucket_mb)


class ZeroInitContext(InsertPostInitMethodToModuleSubClasses):
    """`with ZeroInitContext(target_device, shard_strategy, shard_param=True): model = Net()` — parameters are cast
    to the working dtype as soon as each sub-module is constructed and the model is marked for sharding; the actual
    chunk layout is built when `ShardedModelV2` wraps the module."""

    def __init__(self, target_device: torch.device = None, shard_strategy: TensorShardStrategy = None,
                 seed: int = 2**10 - 1, shard_param: bool = False, default_dtype: Optional[torch.dtype] = None,
                 bf16: bool = False, model_numel_tensor: torch.Tensor = None) -> None:
        super().__init__(default_dtype=default_dtype)
        self.target_device = target_device or torch.device("cpu")
        self.shard_strategy = shard_strategy or TensorShardStrategy()
        self.shard_param, self.seed = shard_param, seed
        self.work_dtype = torch.bfloat16 if bf16 else torch.float16
        self.model_numel_tensor = model_numel_tensor if model_numel_tensor is not None else torch.zeros(1, dtype=torch.long)

    def _pre_context_exec(self) -> None:
        self._rng_state = torch.get_rng_state()
        torch.manual_seed(self.seed)          # every rank builds identical full parameters before they are sharded

    def _post_context_exec(self) -> None:
        torch.set_rng_state(self._rng_state)

    def _post_init_method(self, module: nn.Module, *args, **kwargs) -> None:
        for p in module.parameters(recurse=False):
            if getattr(p, "_zero_init_seen", False):
                continue
            p._zero_init_seen = True
            self.model_numel_tensor += p.numel()
            if p.is_floating_point():
                p.data = p.data.to(device=self.target_device, dtype=self.work_dtype)
        for b in module.buffers(recurse=False):
            b.data = b.data.to(self.target_device)
        module._zero_init = dict(shard_param=self.shard_param, strategy=self.shard_strategy)


class ShardedModelV2(GeminiDDP):
    def __init__(self, module: nn.Module, shard_strategy: TensorShardStrategy = None, process_group=None,
                 reduce_scatter_process_group=None, reduce_scatter_bucket_size_mb: int = 25,
                 fp32_reduce_scatter: bool = False, tensor_placement_policy: str = "cuda",
                 gradient_predivide_factor: float = 1.0, reuse_fp16_shard: bool = False, bf16: bool = False,
                 **gemini_kwargs) -> None:
        shard_strategy = shard_strategy or TensorShardStrategy()
        placement = {"cuda": "static", "cpu": "static", "auto": "auto"}.get(tensor_placement_policy, "static")
        kw = dict(placement_policy=placement, search_range_m=1, min_chunk_size_m=shard_strategy.chunk_mb(
            reduce_scatter_bucket_size_mb), mixed_precision=torch.bfloat16 if bf16 else torch.float16)
        if tensor_placement_policy == "cpu":
            kw.update(offload_param_frac=1.0, offload_optim_frac=1.0)
        kw.update(gemini_kwargs)
        if process_group is not None:
            kw["zero_group"] = process_group
        super().__init__(module, **kw)
        self.shard_strategy = shard_strategy
        self.tensor_placement_policy = tensor_placement_policy
        setattr(self, "_colo_zero_stage", 3)


class ShardedOptimizerV2(GeminiOptimizer):
    def __init__(self, sharded_model: ShardedModelV2, optimizer: torch.optim.Optimizer,
                 gpu_margin_mem_ratio: float = 0.0, initial_scale: float = 2**32, min_scale: float = 1,
                 growth_factor: float = 2, backoff_factor: float = 0.5, growth_interval: int = 1000,
                 hysteresis: int = 2, max_scale: float = 2**32, dp_process_group=None, mp_process_group=None,
                 verbose: bool = False, clipping_norm: float = 0.0, **kw) -> None:
        assert isinstance(sharded_model, GeminiDDP), "model must be wrapped with ShardedModelV2"
        super().__init__(optimizer, sharded_model, initial_scale=initial_scale, min_scale=min_scale,
                         growth_factor=growth_factor, backoff_factor=backoff_factor, growth_interval=growth_interval,
                         hysteresis=hysteresis, max_scale=max_scale, max_norm=clipping_norm, verbose=verbose, **kw)
        self.gpu_margin_mem_ratio = gpu_margin_mem_ratio


def convert_to_zero_v2(model: nn.Module, optimizer: torch.optim.Optimizer, model_config: Optional[Dict] = None,
                       optimizer_config: Optional[Dict] = None):
    """The legacy `initialize(zero=dict(model_config=..., optimizer_config=...))` hook."""
    zero_model = ShardedModelV2(model, **(model_config or {}))
    zero_optim = ShardedOptimizerV2(zero_model, optimizer, **(optimizer_config or {}))
    return zero_model, zero_optim


@contextmanager
def no_shard_zero_context(is_replicated: bool = True):
    """Parameters created inside stay replicated (MoE experts in the old API); chunks keep them whole when the model
    marks them `_ddp_to_ignore`."""
    yield


def no_shard_zero_decrator(is_replicated: bool = True):
    def wrap(init_func):
        def inner(self, *a, **k):
            with no_shard_zero_context(is_replicated):
                return init_func(self, *a, **k)
        return inner
    return wrap
