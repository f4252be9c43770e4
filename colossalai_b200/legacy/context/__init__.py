"""Global parallel context.  Parity: reference `colossalai/legacy/context/{parallel_context,parallel_mode}.py`
(`gpc.get_group(ParallelMode.X)`, `get_local_rank`, `get_world_size`, `is_first_rank`...)."""
from __future__ import annotations

from enum import Enum
from typing import Dict, Optional

import torch.distributed as dist

from ...cluster import DeviceMesh
from ...context import SingletonMeta

__all__ = ["ParallelMode", "ParallelContext", "global_context"]


class ParallelMode(Enum):
    GLOBAL = "global"
    DATA = "data"
    PIPELINE = "pipe"
    TENSOR = "tensor"
    SEQUENCE = "sequence"
    PARALLEL_1D = "1d"
    PARALLEL_2D_ROW = "2d_row"
    PARALLEL_2D_COL = "2d_col"
    PARALLEL_2P5D_ROW = "2p5d_row"
    PARALLEL_2P5D_COL = "2p5d_col"
    PARALLEL_2P5D_DEP = "2p5d_dep"
    PARALLEL_3D_INPUT = "3d_input"
    PARALLEL_3D_WEIGHT = "3d_weight"
    PARALLEL_3D_OUTPUT = "3d_output"


class ParallelContext(metaclass=SingletonMeta):
    """Registry {ParallelMode: process group}; tensor-parallel meshes are created by `init_tensor_mesh`."""

    def __init__(self) -> None:
        self._groups: Dict[ParallelMode, Optional[dist.ProcessGroup]] = {}
        self.mesh: Optional[DeviceMesh] = None
        self.tensor_mode: Optional[str] = None
        self.tensor_dims: Dict[str, int] = {}

    def set_group(self, mode: ParallelMode, group) -> None:
        self._groups[mode] = group

    def get_group(self, mode: ParallelMode):
        if mode == ParallelMode.GLOBAL:
            return dist.group.WORLD if dist.is_initialized() else None
        return self._groups[mode]

    def is_initialized(self, mode: ParallelMode) -> bool:
        return mode == ParallelMode.GLOBAL or mode in self._groups

    def get_world_size(self, mode: ParallelMode) -> int:
        g = self.get_group(mode)
        return dist.get_world_size(g) if dist.is_initialized() else 1

    def get_local_rank(self, mode: ParallelMode) -> int:
        g = self.get_group(mode)
        return dist.get_rank(g) if dist.is_initialized() else 0

    def get_global_rank(self) -> int:
        return dist.get_rank() if dist.is_initialized() else 0

    def is_first_rank(self, mode: ParallelMode) -> bool:
        return self.get_local_rank(mode) == 0

    def is_last_rank(self, mode: ParallelMode) -> bool:
        return self.get_local_rank(mode) == self.get_world_size(mode) - 1

    # ---- tensor-parallel meshes
    def init_tensor_mesh(self, mode: str, size: int, depth: int = 1) -> None:
        """mode: '2d' (size = q*q), '2.5d' (size = d*q*q), '3d' (size = q*q*q).  Tensor ranks are assumed contiguous
        (world = dp x size, tensor innermost)."""
        world = dist.get_world_size()
        assert world % size == 0
        self.tensor_mode = mode
        if mode == "2d":
            q = int(round(size ** 0.5))
            assert q * q == size, "2D tensor parallelism needs a square number of ranks"
            self.mesh = DeviceMesh(dp=world // size, row=q, col=q)
            # ROW group = ranks of the same row (varying column) and vice versa
            self.set_group(ParallelMode.PARALLEL_2D_ROW, self.mesh.group("col"))
            self.set_group(ParallelMode.PARALLEL_2D_COL, self.mesh.group("row"))
            self.tensor_dims = {"q": q}
        elif mode == "2.5d":
            q = int(round((size // depth) ** 0.5))
            assert q * q * depth == size, "2.5D tensor parallelism needs depth * q * q ranks"
            self.mesh = DeviceMesh(dp=world // size, dep=depth, row=q, col=q)
            self.set_group(ParallelMode.PARALLEL_2P5D_ROW, self.mesh.group("col"))
            self.set_group(ParallelMode.PARALLEL_2P5D_COL, self.mesh.group("row"))
            self.set_group(ParallelMode.PARALLEL_2P5D_DEP, self.mesh.group("dep"))
            self.tensor_dims = {"q": q, "d": depth}
        elif mode == "3d":
            q = int(round(size ** (1.0 / 3)))
            assert q ** 3 == size, "3D tensor parallelism needs a cubic number of ranks"
            self.mesh = DeviceMesh(dp=world // size, i=q, j=q, k=q)
            self.set_group(ParallelMode.PARALLEL_3D_INPUT, self.mesh.group("j"))
            self.set_group(ParallelMode.PARALLEL_3D_WEIGHT, self.mesh.group("i"))
            self.set_group(ParallelMode.PARALLEL_3D_OUTPUT, self.mesh.group("k"))
            self.tensor_dims = {"q": q}
        else:
            raise ValueError(f"unknown tensor parallel mode {mode}")
        self.set_group(ParallelMode.DATA, self.mesh.group("dp"))


global_context = ParallelContext()
