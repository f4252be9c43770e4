"""Legacy bring-up: `launch(config, ...)` creates the global parallel context from a config's `parallel` section and
`initialize(model, optimizer, criterion, ...)` returns an `Engine` (+ dataloaders, lr scheduler) with AMP, ZeRO,
gradient handlers and gradient clipping applied as the config asks.

Parity: reference `colossalai/legacy/initialize.py:1-470` (`launch*`, `initialize`), `legacy/global_variables.py`,
`legacy/context/parallel_context.py:init_parallel_groups`."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, Dict, Iterable, Optional, Tuple, Union

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.optim import Optimizer
from torch.utils.data import DataLoader

from ..cluster import DeviceMesh
from ..context import Config
from ..initialize import launch as _launch
from .amp import AMP_TYPE, convert_to_amp
from .context import ParallelMode, global_context as gpc
from .engine import DataParallelGradientHandler, Engine

__all__ = ["launch", "launch_from_torch", "initialize", "get_default_parser"]


def get_default_parser():
    import argparse

    p = argparse.ArgumentParser()
    p.add_argument("--config", type=str, help="path to the config file")
    p.add_argument("--host", type=str, help="the master address for distributed training")
    p.add_argument("--port", type=int, help="the master port for distributed training")
    p.add_argument("--world_size", type=int, help="world size for distributed training")
    p.add_argument("--rank", type=int, help="rank for the default process group")
    p.add_argument("--local_rank", type=int, help="local rank on the node")
    p.add_argument("--backend", type=str, default="nccl", help="backend for distributed communication")
    return p


def _load_config(config: Union[str, Path, Config, Dict, None]) -> Config:
    if config is None:
        return Config()
    if isinstance(config, (str, Path)):
        return Config.from_file(str(config))
    return config if isinstance(config, Config) else Config(config)


def _init_parallel_groups(cfg: Config) -> None:
    """`parallel = dict(pipeline=P, tensor=dict(size=T, mode='1d'|'2d'|'2.5d'|'3d'|'sequence', depth=D))`."""
    world = dist.get_world_size()
    par = cfg.get("parallel", {}) or {}
    pp = par.get("pipeline", 1)
    pp = pp.get("size", 1) if isinstance(pp, dict) else int(pp)
    tcfg = par.get("tensor", {}) or {}
    tp = int(tcfg.get("size", 1)) if isinstance(tcfg, dict) else int(tcfg)
    mode = (tcfg.get("mode") if isinstance(tcfg, dict) else None) or "1d"
    assert world % (pp * tp) == 0, f"world size {world} not divisible by pipeline {pp} x tensor {tp}"
    dp = world // (pp * tp)
    mesh = DeviceMesh(dp=dp, pp=pp, tp=tp)
    gpc.mesh = mesh
    gpc.set_group(ParallelMode.DATA, mesh.group("dp"))
    gpc.set_group(ParallelMode.PIPELINE, mesh.group("pp"))
    gpc.set_group(ParallelMode.TENSOR, mesh.group("tp"))
    if mode == "1d":
        gpc.set_group(ParallelMode.PARALLEL_1D, mesh.group("tp"))
    elif mode == "sequence":
        gpc.set_group(ParallelMode.SEQUENCE, mesh.group("tp"))
    elif tp > 1:
        gpc.init_tensor_mesh(mode, tp, depth=int(tcfg.get("depth", 1)))
    gpc.tensor_mode = mode
    gpc.config = cfg


def launch(config=None, rank: int = 0, world_size: int = 1, host: str = "127.0.0.1", port: int = 29500,
           backend: str = "nccl", local_rank: Optional[int] = None, seed: int = 1024, verbose: bool = True) -> None:
    cfg = _load_config(config)
    _launch(rank, world_size, host, port, backend=backend, local_rank=local_rank, seed=seed, verbose=verbose)
    _init_parallel_groups(cfg)


def launch_from_torch(config=None, backend: str = "nccl", seed: int = 1024, verbose: bool = True) -> None:
    launch(config, rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]),
           host=os.environ.get("MASTER_ADDR", "127.0.0.1"), port=int(os.environ.get("MASTER_PORT", 29500)),
           backend=backend, local_rank=int(os.environ.get("LOCAL_RANK", 0)), seed=seed, verbose=verbose)


def initialize(model: nn.Module, optimizer: Optimizer, criterion: Optional[Callable] = None,
               train_dataloader: Optional[Iterable] = None, test_dataloader: Optional[Iterable] = None,
               lr_scheduler=None, ophooks=None, verbose: bool = True) -> Tuple[Engine, DataLoader, DataLoader, object]:
    cfg: Config = getattr(gpc, "config", None) or Config()
    if callable(model) and not isinstance(model, nn.Module):
        model = model()
    if callable(optimizer) and not isinstance(optimizer, Optimizer):
        optimizer = optimizer(model.parameters())
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    zero_cfg = cfg.get("zero", None)
    if zero_cfg is not None:
        from .zero import convert_to_zero_v2

        model, optimizer = convert_to_zero_v2(model, optimizer, zero_cfg.get("model_config"),
                                              zero_cfg.get("optimizer_config"))
    else:
        model = model.to(dev)
        from .utils import is_using_ddp, sync_model_param

        if is_using_ddp():
            sync_model_param(model, ParallelMode.DATA)
    fp16 = cfg.get("fp16", None)
    if fp16 is not None and fp16.get("mode") is not None and zero_cfg is None:
        amp_cfg = {k: v for k, v in fp16.items() if k != "mode"}
        model, optimizer, criterion = convert_to_amp(model, optimizer, criterion, fp16["mode"], amp_cfg)
    handlers = []
    if zero_cfg is None and gpc.is_initialized(ParallelMode.DATA) and gpc.get_world_size(ParallelMode.DATA) > 1:
        handlers.append(DataParallelGradientHandler(model, gpc.get_group(ParallelMode.DATA)))
    engine = Engine(model, optimizer, criterion, gradient_handlers=handlers,
                    clip_grad_norm=float(cfg.get("clip_grad_norm", 0.0)), verbose=verbose)
    return engine, train_dataloader, test_dataloader, lr_scheduler
