"""Static analysis helpers: device-free tensors, FLOP counting and per-node symbolic profiling of fx graphs.

Parity: reference `colossalai/_analyzer` (`_subclasses/{meta_tensor,flop_tensor}.py`, `fx/{symbolic_profile,
node_util,graph_module,tracer}`, `envs.py`)."""
from . import fx
from ._subclasses import MetaTensor, MetaTensorMode, flop_count, flop_mapping
from .envs import MeshConfig

__all__ = ["MetaTensor", "MetaTensorMode", "flop_count", "flop_mapping", "MeshConfig", "fx"]
