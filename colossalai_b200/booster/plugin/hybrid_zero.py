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
                 overlap_allgather: bool = False, fp8_communication: bool = False, offload_optim_frac: float = 1.0,
                 **unused) -> None:
        from .hybrid_parallel_plugin import _reassign_params

        self.model = model
        self.param_info = param_info
        self.stage_manager = model.stage_manager
        self.shared_params = model.shared_params
        self.tp_pg, self.pp_pg = tp_process_group, pp_process_group
        self.tp_size = comm.group_size(tp_process_group) if tp_process_group is not None else 1
        self.pp_size = comm.group_size(pp_process_group) if pp_process_group is not None else 1
        _reassign_params(optimizer, model, param_info)
        super().__init__(optimizer=optimizer, pg_to_param_list=pg_to_param_list, initial_scale=initial_scale,
                         min_scale=min_scale, growth_factor=growth_factor, backoff_factor=backoff_factor,
                         growth_interval=growth_interval, hysteresis=hysteresis, max_scale=max_scale,
                         clip_grad_norm=clip_grad_norm, verbose=verbose, reduce_bucket_size=reduce_bucket_size,
                         communication_dtype=communication_dtype, overlap_communication=overlap_communication,
                         partition_grad=partition_grad, cpu_offload=cpu_offload, dp_process_group=dp_process_group,
                         forced_dtype=forced_dtype, overlap_allgather=overlap_allgather,
                         fp8_communication=fp8_communication, offload_optim_frac=offload_optim_frac)

    def _reduce_bucket(self, b) -> None:
        """The dp reduce-scatter is the ONLY place a bucket's local gradient leaves `grad_full` (the grad hook calls it
        directly for ZeRO-2 / overlapped ZeRO-1, `sync_grad` for the rest), so the two syncs that must precede it live
        here (reference: `hybrid_parallel_plugin.py:88-137` + `layer/utils.py:75-127`, which run on `p.grad` before
        the ZeRO reduction):
          * tied parameters across pipeline stages: sum the tied slices over the shared-param group;
          * split_gather / ring sequence parallelism: norm weights and row-linear biases only saw a sequence slice,
            sum their slices over the tp group.
        Every rank of a tp (resp. shared) group walks identical buckets in identical order, so the collectives match."""
        if b.grad_full is None:
            return
        tied = self._tied_params()
        if tied:
            for p, o in zip(b.params, b.offsets):
                group = tied.get(id(p))
                if group is not None:
                    dist.all_reduce(b.grad_full[o:o + p.numel()], group=group)
        sc = self.model.shard_config
        if sc.sp_mode in ("split_gather", "ring") and self.tp_size > 1:
            slices = [b.grad_full[o:o + p.numel()] for p, o in zip(b.params, b.offsets)
                      if getattr(p, "partial_derived", False)]
            if len(slices) == 1:
                dist.all_reduce(slices[0], group=self.tp_pg)
            elif slices:
                flat = torch.cat(slices)
                dist.all_reduce(flat, group=self.tp_pg)
                off = 0
                for s in slices:
                    s.copy_(flat[off:off + s.numel()])
                    off += s.numel()
        super()._reduce_bucket(b)

    def _tied_params(self) -> Dict[int, ProcessGroup]:
        """id(param) -> process group of the pipeline stages that hold a copy of the same tied weight."""
        cache = getattr(self, "_tied_cache", None)
        if cache is None:
            cache = {}
            sm = self.stage_manager
            groups = getattr(self.model, "shared_param_process_groups", [])
            if sm is not None:
                for shared, group in zip([s for s in self.shared_params if len(s) > 0], groups):
                    if sm.stage in shared and comm.group_size(group) > 1:
                        cache[id(shared[sm.stage])] = group
            self._tied_cache = cache
        return cache

    def _compute_grad_norm_sq(self) -> Tensor:
        """Bucket shards mix TP-sharded, replicated and expert-parallel params and may live on different dp groups
        (dense: dp, experts: moe_dp): accumulate [tp-sharded, replicated, moe] per group, reduce each over its group."""
        from ...tensor.moe_tensor import get_ep_group, is_moe_tensor

        dev = self.buckets[0].device if self.buckets else torch.device("cpu")
        shared_ids = set()
        if self.stage_manager is not None:
            for shared in self.shared_params:
                for s in sorted(shared.keys())[1:]:
                    shared_ids.add(id(shared[s]))
        per_pg = {}
        ep_group = None
        for b in self.buckets:
            if b.grad_shard is None:
                continue
            acc = per_pg.setdefault(id(b.pg), (b.pg, torch.zeros(3, device=dev)))[1]
            lo, hi = b.my_slice.start, b.my_slice.stop
            for p, o in zip(b.params, b.offsets):
                s, e = max(o, lo), min(o + p.numel(), hi)
                if e <= s or id(p) in shared_ids:
                    continue
                v = b.grad_shard[s - lo:e - lo].float().pow(2).sum()
                if is_moe_tensor(p):
                    acc[2] += v
                    ep_group = get_ep_group(p)
                elif self.tp_size > 1 and is_distributed_tensor(p):
                    acc[0] += v
                else:
                    acc[1] += v
        total3 = torch.zeros(3, device=dev)
        for pg, acc in per_pg.values():
            if comm.group_size(pg) > 1:      # every rank of the group holds a different slice
                dist.all_reduce(acc, group=pg)
            total3 += acc
        sharded, replicated, moe = total3[0:1], total3[1:2], total3[2:3]
        if self.tp_size > 1:
            dist.all_reduce(sharded, group=self.tp_pg)
        if getattr(self, "ep_pg", None) is not None:
            ep_group = self.ep_pg
        if ep_group is not None and comm.group_size(ep_group) > 1:
            dist.all_reduce(moe, group=ep_group)       # collective: every rank of the ep group calls it
        total = sharded + replicated + moe
        if self.pp_size > 1:
            dist.all_reduce(total, group=self.pp_pg)
        return total

    def get_working_to_master_map(self):
        return super().get_working_to_master_map()
