"""ZeRO-1/2 optimizer for hybrid-parallel models (grad norm reduced over dp, tp and pp).
Parity: reference `HybridParallelZeroOptimizer` (`colossalai/booster/plugin/hybrid_parallel_plugin.py:666-926`)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor
from torch.distributed import ProcessGroup
from torch.optim import Optimizer

from ...parallel import comm
from ...tensor.d_tensor import is_distributed_tensor
from ...tensor.moe_tensor import is_moe_tensor
from ...zero.low_level import LowLevelZeroOptimizer

__all__ = ["HybridParallelZeroOptimizer"]


class HybridParallelZeroOptimizer(LowLevelZeroOptimizer):
    def __init__(self, optimizer: Optimizer, model, use_pipeline: bool, param_info: Dict,
                 pg_to_param_list: Optional[Dict] = None, initial_scale: float = 2**16, min_scale: float = 1,
                 growth_factor: float = 2.0, backoff_factor: float = 0.5, growth_interval: int = 2000,
                 hysteresis: int = 2, max_scale: float = 2**24, clip_grad_norm: float = 0.0, verbose: bool = False,
                 reduce_bucket_size: int = 1024 * 1024, communication_dtype: Optional[torch.dtype] = None,
                 overlap_communication: bool = True, partition_grad: bool = False, cpu_offload: bool = False,
                 dp_process_group: Optional[ProcessGroup] = None, tp_process_group: Optional[ProcessGroup] = None,
                 pp_process_group: Optional[ProcessGroup] = None, forced_dtype: Optional[torch.dtype] = None,
                 overlap_allgather: bool = False, fp8_communication: bool = False, **unused) -> None:
        from .hybrid_parallel_plugin import _reassign_params

        self.model = model
        self.param_info = param_info
        self.stage_manager = model.stage_manager
        self.shared_params = model.shared_params
        self.tp_pg, self.pp_pg = tp_process_group, pp_process_group
        self.tp_size = comm.group_size(tp_process_group) if tp_process_group is not None else 1
        self.pp_size = comm.group_size(pp_process_group) if pp_process_group is not None else 1
        _reassign_params(optimizer, model)
        super().__init__(optimizer=optimizer, pg_to_param_list=pg_to_param_list, initial_scale=initial_scale,
                         min_scale=min_scale, growth_factor=growth_factor, backoff_factor=backoff_factor,
                         growth_interval=growth_interval, hysteresis=hysteresis, max_scale=max_scale,
                         clip_grad_norm=clip_grad_norm, verbose=verbose, reduce_bucket_size=reduce_bucket_size,
                         communication_dtype=communication_dtype, overlap_communication=overlap_communication,
                         partition_grad=partition_grad, cpu_offload=cpu_offload, dp_process_group=dp_process_group,
                         forced_dtype=forced_dtype, overlap_allgather=overlap_allgather,
                         fp8_communication=fp8_communication)

    def sync_grad(self) -> None:
        # SP-partial grads live inside the flat bucket buffers: all-reduce those param slices over tp first
        sc = self.model.shard_config
        if sc.sp_mode in ("split_gather", "ring") and self.tp_size > 1:
            for b in self.buckets:
                if b.grad_full is None:
                    continue
                for p, o in zip(b.params, b.offsets):
                    if getattr(p, "partial_derived", False):
                        dist.all_reduce(b.grad_full[o:o + p.numel()], group=self.tp_pg)
        super().sync_grad()

    def _compute_grad_norm_sq(self) -> Tensor:
        """Bucket shards mix TP-sharded and replicated params: split the local sum accordingly."""
        dev = self.buckets[0].device if self.buckets else torch.device("cpu")
        sharded = torch.zeros(1, device=dev)
        replicated = torch.zeros(1, device=dev)
        shared_ids = set()
        if self.stage_manager is not None:
            for shared in self.shared_params:
                for s in sorted(shared.keys())[1:]:
                    shared_ids.add(id(shared[s]))
        for b in self.buckets:
            if b.grad_shard is None:
                continue
            lo, hi = b.my_slice.start, b.my_slice.stop
            for p, o in zip(b.params, b.offsets):
                s, e = max(o, lo), min(o + p.numel(), hi)
                if e <= s or id(p) in shared_ids:
                    continue
                v = b.grad_shard[s - lo:e - lo].float().pow(2).sum().reshape(1)
                if self.tp_size > 1 and is_distributed_tensor(p):
                    sharded += v
                else:
                    replicated += v
        # sum over the dp group(s): every dp rank holds a different slice
        by_pg = {}
        for b in self.buckets:
            by_pg.setdefault(id(b.pg), b.pg)
        both = torch.cat([sharded, replicated])
        for pg in by_pg.values():
            if comm.group_size(pg) > 1:
                dist.all_reduce(both, group=pg)
                break
        sharded, replicated = both[0:1], both[1:2]
        if self.tp_size > 1:
            dist.all_reduce(sharded, group=self.tp_pg)
        total = sharded + replicated
        if self.pp_size > 1:
            dist.all_reduce(total, group=self.pp_pg)
        return total

    def get_working_to_master_map(self):
        return super().get_working_to_master_map()
