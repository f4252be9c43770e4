"""Named registries of meshes / process groups.  Parity: reference `cluster/device_mesh_manager.py:58`,
`cluster/process_group_manager.py:7`."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch.distributed as dist
from torch.distributed import ProcessGroup

__all__ = ["ProcessGroupManager", "DeviceMeshManager"]


class ProcessGroupManager:
    def __init__(self) -> None:
        self.pg_store: Dict[str, ProcessGroup] = {}

    def create_process_group(self, name: str, ranks: List[int], backend: Optional[str] = None) -> ProcessGroup:
        if name in self.pg_store:
            raise ValueError(f"process group {name!r} already exists")
        pg = dist.new_group(ranks=ranks, backend=backend)
        self.pg_store[name] = pg
        return pg

    def get(self, name: str) -> ProcessGroup:
        return self.pg_store[name]

    def destroy(self, name: str) -> None:
        pg = self.pg_store.pop(name)
        dist.destroy_process_group(pg)

    def destroy_all(self) -> None:
        for name in list(self.pg_store):
            self.destroy(name)


class DeviceMeshManager:
    def __init__(self) -> None:
        self.device_mesh_store: Dict[str, object] = {}

    def create_device_mesh(self, name: str, **axes: int):
        from .mesh import DeviceMesh

        if name in self.device_mesh_store:
            raise ValueError(f"device mesh {name!r} already exists")
        mesh = DeviceMesh(**axes)
        self.device_mesh_store[name] = mesh
        return mesh

    def get(self, name: str):
        return self.device_mesh_store[name]

    def destroy(self, name: str) -> None:
        mesh = self.device_mesh_store.pop(name)
        mesh.destroy_mesh_process_groups()

    def destroy_all(self) -> None:
        for name in list(self.device_mesh_store):
            self.destroy(name)
