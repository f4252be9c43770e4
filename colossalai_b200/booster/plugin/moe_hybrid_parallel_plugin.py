"""MoeHybridParallelPlugin: HybridParallelPlugin + expert parallelism.

Parity: reference `colossalai/booster/plugin/moe_hybrid_parallel_plugin.py:53-520` (`MoeHybridParallelZeroOptimizer`
with separate dp groups for dense / expert params, `ep_size` carved out of dp, `moe_dp_group`, `MoECheckpointIO`).

Mesh: the data-parallel axis of the parent plugin is factorised as dp = moe_dp x ep.  Dense parameters are replicated
over dp (gradients averaged over dp [x sp]); expert parameters are sharded over ep and replicated over moe_dp.  An expert
sees the tokens of every ep rank, so its local gradient is the SUM over ep micro-batches: it is scaled by 1/ep (hook) and
averaged over moe_dp, which makes it the mean over the global batch like every dense gradient.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler as LRScheduler
from torch.utils.data import DataLoader

from ...checkpoint_io import CheckpointIO
from ...cluster import DeviceMesh
from ...interface import ModelWrapper, OptimizerWrapper
from ...parallel import comm
from ...tensor.moe_tensor import is_moe_tensor
from .hybrid_parallel_plugin import HybridParallelModule, HybridParallelPlugin

__all__ = ["MoeHybridParallelPlugin", "MoeHybridParallelZeroOptimizer"]


def _scale_hook(scale: float):
    """Tensor-level hook: scales only the INCOMING gradient of one backward.  (A post-accumulate hook that rescales
    `p.grad` would rescale the contributions of earlier micro-batches again on every backward.)"""
    def hook(grad: torch.Tensor) -> torch.Tensor:
        return grad * scale

    return hook


class MoeHybridParallelPlugin(HybridParallelPlugin):
    """
    >>> plugin = MoeHybridParallelPlugin(ep_size=8, tp_size=1, pp_size=1, zero_stage=1, precision="bf16")
    >>> booster = Booster(plugin=plugin)
    """

    def __init__(self, tp_size: int = 1, pp_size: int = 1, ep_size: int = 1, moe_dp_outside: bool = True,
                 overlap_p2p: bool = True, **kwargs) -> None:
        if kwargs.get("overlap_communication") and kwargs.get("zero_stage", 0) > 0 and ep_size > 1:
            # expert tokens arrive late in the backward: overlapping the bucket reduce with them buys nothing
            kwargs["overlap_communication"] = False
        super().__init__(tp_size=tp_size, pp_size=pp_size, overlap_p2p=overlap_p2p, **kwargs)
        assert self.dp_size % ep_size == 0, f"dp size {self.dp_size} must be divisible by ep size {ep_size}"
        self.ep_size = ep_size
        self.moe_dp_size = self.dp_size // ep_size
        world = dist.get_world_size()
        if ep_size == 1:
            # degenerate: experts are plain data-parallel parameters
            self.ep_group = self._self_group()
            self.moe_dp_group = self.dp_group
            self.moe_mesh = None
        else:
            # same rank order as the parent mesh with the dp axis split in two
            order = [a for a in self.pg_mesh.axis_names]
            sizes = {a: self.pg_mesh.axis_size(a) for a in order}
            new_axes = {}
            for a in order:
                if a == "dp":
                    if moe_dp_outside:
                        new_axes["moe_dp"], new_axes["ep"] = self.moe_dp_size, ep_size
                    else:
                        new_axes["ep"], new_axes["moe_dp"] = ep_size, self.moe_dp_size
                else:
                    new_axes[a] = sizes[a]
            self.moe_mesh = DeviceMesh(**new_axes)
            self.ep_group = self.moe_mesh.group("ep")
            self.moe_dp_group = self.moe_mesh.group("moe_dp")
        self.shard_config.ep_group = self.ep_group
        self.shard_config.moe_dp_group = self.moe_dp_group
        self.shard_config._expert_parallel_size = ep_size

    def _self_group(self):
        """One single-rank group per rank (collective creation: every rank creates all of them)."""
        mine = None
        for r in range(dist.get_world_size()):
            g = dist.new_group([r])
            if r == dist.get_rank():
                mine = g
        return mine

    def configure(self, model: nn.Module, optimizer: Optional[Optimizer] = None, criterion: Optional[Callable] = None,
                  dataloader: Optional[DataLoader] = None, lr_scheduler: Optional[LRScheduler] = None):
        from .hybrid_parallel_plugin import get_param_info
        from .hybrid_zero import HybridParallelZeroOptimizer

        if self.zero_stage == 0 or optimizer is None or isinstance(optimizer, OptimizerWrapper):
            model, optimizer, criterion, dataloader, lr_scheduler = super().configure(model, optimizer, criterion,
                                                                                      dataloader, lr_scheduler)
            self._tag(model)
            return model, optimizer, criterion, dataloader, lr_scheduler
        # ZeRO: dense params are partitioned over dp (x sp), expert params over moe_dp
        param_info = get_param_info(optimizer, model)
        if not isinstance(model, ModelWrapper):
            model = HybridParallelModule(model, precision=self.precision, shard_config=self.shard_config,
                                         dp_group=self.mixed_dp_group, tp_group=self.tp_group, sp_group=self.sp_group,
                                         use_ddp=False, ddp_config=self.ddp_config, custom_policy=self.custom_policy,
                                         overlap_allgather=self.zero_config["overlap_allgather"], use_fp8=self.use_fp8)
            model.dp_size = self.dp_size_for_grads
        self._tag(model)
        from ...nn.optimizer import cast_to_distributed

        optimizer = cast_to_distributed(optimizer)
        from .hybrid_parallel_plugin import _reassign_params

        _reassign_params(optimizer, model, param_info)
        dense = [p for g in optimizer.param_groups for p in g["params"] if not is_moe_tensor(p)]
        moe = [p for g in optimizer.param_groups for p in g["params"] if is_moe_tensor(p)]
        pg_map = {self.mixed_dp_group: dense}
        if moe:
            pg_map[self.moe_dp_group] = moe
        assert self.precision != "fp32", "Please set precision to 'fp16' or 'bf16' when using ZeRO."
        optimizer = MoeHybridParallelZeroOptimizer(
            optimizer, model, use_pipeline=self.enable_pipeline_parallelism, param_info=param_info,
            pg_to_param_list=pg_map, dp_process_group=self.mixed_dp_group, tp_process_group=self.tp_group,
            pp_process_group=self.pp_group, verbose=False, clip_grad_norm=self.max_norm, **self.zero_config,
            **self.amp_config)
        optimizer.ep_pg = self.ep_group
        if hasattr(model, "bind_optimizer"):
            model.bind_optimizer(optimizer)
        return model, optimizer, criterion, dataloader, lr_scheduler

    def _tag(self, model: ModelWrapper) -> None:
        """moe_dp sync group on the wrapper + the 1/ep gradient scaling on expert parameters."""
        model.moe_dp_group = self.moe_dp_group
        if getattr(model, "_moe_hooks_done", False):
            return
        model._moe_hooks_done = True
        if self.ep_size > 1:
            for p in model.unwrap().parameters():
                if is_moe_tensor(p) and p.requires_grad:
                    p.register_hook(_scale_hook(1.0 / self.ep_size))

    def get_checkpoint_io(self) -> CheckpointIO:
        from ...checkpoint_io import MoECheckpointIO

        return MoECheckpointIO(self.mixed_dp_group, self.pp_group, self.tp_group, self.sp_group, self.ep_group,
                               self.moe_dp_group, self.zero_stage)


from .hybrid_zero import HybridParallelZeroOptimizer  # noqa: E402


class MoeHybridParallelZeroOptimizer(HybridParallelZeroOptimizer):
    """ZeRO over two data-parallel groups (dense -> dp, experts -> moe_dp); the norm reduction over ep is handled by
    `HybridParallelZeroOptimizer._compute_grad_norm_sq` through `ep_pg`."""

    ep_pg = None
