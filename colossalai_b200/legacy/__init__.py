"""Legacy API surface (reference `colossalai/legacy`): global parallel context, 2D / 2.5D / 3D tensor-parallel layers,
engine / trainer with hooks.  Kept small and built on the current runtime (DeviceMesh, comm)."""
from .context import ParallelContext, ParallelMode, global_context

__all__ = ["ParallelContext", "ParallelMode", "global_context"]
