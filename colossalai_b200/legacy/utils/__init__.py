"""Legacy helper grab-bag.  Parity: reference `colossalai/legacy/utils/{common.py:1-430, activation_checkpoint.py:1-270,
data_sampler/data_parallel_sampler.py:1-160, memory.py, checkpointing.py}`."""
from __future__ import annotations

import math
import random
from typing import Iterator, Optional, TypeVar

import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset, Sampler

from ..context import ParallelMode, global_context as gpc

T_co = TypeVar("T_co", covariant=True)

__all__ = ["checkpoint", "DataParallelSampler", "get_dataloader", "sync_model_param", "clip_grad_norm_fp32",
           "count_zeros_fp32", "is_using_ddp", "is_using_pp", "is_using_sequence", "is_dp_rank_0", "is_tp_rank_0",
           "print_rank_0", "report_memory_usage", "is_no_pp_or_last_stage", "conditional_context"]


# ------------------------------------------------------------------------------------- activation checkpoint
class _CheckpointFunction(torch.autograd.Function):
    """Recompute-in-backward with RNG state restore and optional host offload of the saved inputs."""

    @staticmethod
    def forward(ctx, run_function, activation_offload, *args):
        ctx.run_function, ctx.activation_offload = run_function, activation_offload
        ctx.cpu_rng = torch.get_rng_state()
        ctx.cuda_rng = torch.cuda.get_rng_state() if torch.cuda.is_available() else None
        ctx.tensor_idx, ctx.inputs, tensors = [], [], []
        for i, a in enumerate(args):
            if torch.is_tensor(a):
                ctx.tensor_idx.append(i)
                ctx.inputs.append(None)
                t = a.detach()
                if activation_offload:
                    ctx.dev = a.device
                    t = t.to("cpu", non_blocking=True)
                tensors.append(t)
            else:
                ctx.inputs.append(a)
        ctx.req = [torch.is_tensor(a) and a.requires_grad for a in args]
        ctx.save_for_backward(*tensors)
        with torch.no_grad():
            return run_function(*args)

    @staticmethod
    def backward(ctx, *grads):
        inputs = list(ctx.inputs)
        for k, i in enumerate(ctx.tensor_idx):
            t = ctx.saved_tensors[k]
            if ctx.activation_offload:
                t = t.to(ctx.dev)
            inputs[i] = t.detach().requires_grad_(ctx.req[i])
        cpu_now = torch.get_rng_state()
        cuda_now = torch.cuda.get_rng_state() if ctx.cuda_rng is not None else None
        torch.set_rng_state(ctx.cpu_rng)
        if ctx.cuda_rng is not None:
            torch.cuda.set_rng_state(ctx.cuda_rng)
        with torch.enable_grad():
            out = ctx.run_function(*inputs)
        torch.set_rng_state(cpu_now)
        if cuda_now is not None:
            torch.cuda.set_rng_state(cuda_now)
        outs = (out,) if torch.is_tensor(out) else tuple(out)
        pairs = [(o, g) for o, g in zip(outs, grads) if torch.is_tensor(o) and o.requires_grad and g is not None]
        torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        return (None, None) + tuple(x.grad if torch.is_tensor(x) and x.requires_grad else None for x in inputs)


def checkpoint(function, activation_offload: bool, *args, use_reentrant: bool = True):
    """`checkpoint(fn, offload, *inputs)`: recompute `fn` in backward; with `activation_offload` the saved inputs wait
    on the host in between."""
    if use_reentrant or activation_offload:
        return _CheckpointFunction.apply(function, activation_offload, *args)
    from torch.utils.checkpoint import checkpoint as tc

    return tc(function, *args, use_reentrant=False)


# ------------------------------------------------------------------------------------- data
class DataParallelSampler(Sampler):
    """Shards a dataset over the data-parallel group (seeded shuffle, optional drop_last / padding)."""

    def __init__(self, dataset: Dataset, shuffle: bool = False, seed: int = 0, drop_last: bool = False) -> None:
        self.dataset = dataset
        self.num_replicas = gpc.get_world_size(ParallelMode.DATA) if gpc.is_initialized(ParallelMode.DATA) else (
            dist.get_world_size() if dist.is_initialized() else 1)
        self.rank = gpc.get_local_rank(ParallelMode.DATA) if gpc.is_initialized(ParallelMode.DATA) else (
            dist.get_rank() if dist.is_initialized() else 0)
        self.epoch, self.shuffle, self.seed, self.drop_last = 0, shuffle, seed, drop_last
        n = len(dataset)
        self.num_samples = (n // self.num_replicas) if (drop_last and n % self.num_replicas) else math.ceil(
            n / self.num_replicas)
        self.total_size = self.num_samples * self.num_replicas

    def __iter__(self) -> Iterator[T_co]:
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            idx = torch.randperm(len(self.dataset), generator=g).tolist()
            self.epoch += 1
        else:
            idx = list(range(len(self.dataset)))
        if not self.drop_last:
            pad = self.total_size - len(idx)
            idx += (idx * math.ceil(pad / max(1, len(idx))))[:pad]
        else:
            idx = idx[: self.total_size]
        return iter(idx[self.rank: self.total_size: self.num_replicas])

    def __len__(self) -> int:
        return self.num_samples

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch


def get_dataloader(dataset, shuffle: bool = False, seed: int = 1024, add_sampler: bool = True, drop_last: bool = False,
                   pin_memory: bool = False, num_workers: int = 0, **kwargs) -> DataLoader:
    sampler = DataParallelSampler(dataset, shuffle=shuffle, seed=seed, drop_last=drop_last) if add_sampler else None

    def seed_worker(worker_id):
        np.random.seed(seed)
        torch.manual_seed(seed)
        random.seed(seed)

    return DataLoader(dataset, sampler=sampler, shuffle=(shuffle and sampler is None), worker_init_fn=seed_worker,
                      drop_last=drop_last, pin_memory=pin_memory, num_workers=num_workers, **kwargs)


# ------------------------------------------------------------------------------------- misc
def is_using_ddp() -> bool:
    return gpc.is_initialized(ParallelMode.DATA) and gpc.get_world_size(ParallelMode.DATA) > 1


def is_using_pp() -> bool:
    return gpc.is_initialized(ParallelMode.PIPELINE) and gpc.get_world_size(ParallelMode.PIPELINE) > 1


def is_using_sequence() -> bool:
    return gpc.is_initialized(ParallelMode.SEQUENCE) and gpc.get_world_size(ParallelMode.SEQUENCE) > 1


def is_dp_rank_0() -> bool:
    return not gpc.is_initialized(ParallelMode.DATA) or gpc.is_first_rank(ParallelMode.DATA)


def is_tp_rank_0() -> bool:
    return not gpc.is_initialized(ParallelMode.TENSOR) or gpc.is_first_rank(ParallelMode.TENSOR)


def is_no_pp_or_last_stage() -> bool:
    return not gpc.is_initialized(ParallelMode.PIPELINE) or gpc.is_last_rank(ParallelMode.PIPELINE)


def print_rank_0(msg: str, logger=None) -> None:
    if gpc.get_global_rank() == 0:
        (logger.info if logger is not None else print)(msg)


def sync_model_param(model: torch.nn.Module, parallel_mode: ParallelMode = ParallelMode.DATA) -> None:
    """Broadcast parameters from the first rank of the group so replicas start identical."""
    if gpc.is_initialized(parallel_mode) and gpc.get_world_size(parallel_mode) > 1:
        g = gpc.get_group(parallel_mode)
        src = dist.get_global_rank(g, 0)
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=g)


def clip_grad_norm_fp32(parameters, max_norm: float, norm_type: float = 2.0) -> torch.Tensor:
    """Global-norm clipping where tensor-parallel shards contribute their partial squared norms."""
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return torch.zeros(())
    dev = params[0].grad.device
    if norm_type == math.inf:
        total = torch.stack([p.grad.detach().abs().max() for p in params]).max().float()
        op = dist.ReduceOp.MAX
    else:
        total = torch.stack([p.grad.detach().float().norm(norm_type) ** norm_type for p in params]).sum()
        op = dist.ReduceOp.SUM
    if gpc.is_initialized(ParallelMode.TENSOR) and gpc.get_world_size(ParallelMode.TENSOR) > 1:
        # replicated params would be double counted: only tensor rank 0 adds them
        from ...tensor.d_tensor import is_distributed_tensor

        if norm_type != math.inf and not is_tp_rank_0():
            rep = [p for p in params if not is_distributed_tensor(p)]
            if rep:
                total = total - torch.stack([p.grad.detach().float().norm(norm_type) ** norm_type for p in rep]).sum()
        dist.all_reduce(total, op=op, group=gpc.get_group(ParallelMode.TENSOR))
    if is_using_pp():
        dist.all_reduce(total, op=op, group=gpc.get_group(ParallelMode.PIPELINE))
    norm = total if norm_type == math.inf else total ** (1.0 / norm_type)
    coef = max_norm / (norm + 1e-6)
    if coef < 1.0:
        for p in params:
            p.grad.detach().mul_(coef.to(dev))
    return norm


def count_zeros_fp32(parameters) -> int:
    n = sum(int((p.grad.detach() == 0).sum()) for p in parameters if p.grad is not None)
    t = torch.tensor([n], dtype=torch.long)
    if gpc.is_initialized(ParallelMode.TENSOR) and gpc.get_world_size(ParallelMode.TENSOR) > 1:
        dist.all_reduce(t, group=gpc.get_group(ParallelMode.TENSOR))
    return int(t)


def report_memory_usage(message: str = "", logger=None, report_cpu: bool = False) -> str:
    if torch.cuda.is_available():
        mb = 2 ** 20
        msg = (f"{message} | GPU: allocated {torch.cuda.memory_allocated() / mb:.1f} MB, max allocated "
               f"{torch.cuda.max_memory_allocated() / mb:.1f} MB, reserved {torch.cuda.memory_reserved() / mb:.1f} MB")
        torch.cuda.reset_peak_memory_stats()
    else:
        msg = f"{message} | no CUDA device"
    if report_cpu:
        import psutil

        vm = psutil.virtual_memory()
        msg += f" | CPU: used {vm.used / 2**20:.0f} MB ({vm.percent}%)"
    print_rank_0(msg, logger)
    return msg


class conditional_context:
    def __init__(self, context_manager, enable: bool = True) -> None:
        self.cm, self.enable = context_manager, enable

    def __enter__(self):
        return self.cm.__enter__() if self.enable else None

    def __exit__(self, *exc):
        return self.cm.__exit__(*exc) if self.enable else False
