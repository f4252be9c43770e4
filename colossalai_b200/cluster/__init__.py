from .mesh import DeviceMesh, ProcessGroupMesh
from .dist_coordinator import DistCoordinator
from .registry import DeviceMeshManager, ProcessGroupManager

__all__ = ["DeviceMesh", "ProcessGroupMesh", "DistCoordinator", "DeviceMeshManager", "ProcessGroupManager"]
