"""`ColoTensorSpec(pg, dist_attr, compute_attr)` (reference `legacy/tensor/tensor_spec.py`)."""
from dataclasses import dataclass, field
from typing import Optional

from .compute_spec import ComputeSpec
from .distspec import ReplicaSpec, _DistSpec
from .process_group import ProcessGroup

__all__ = ["ColoTensorSpec"]


@dataclass
class ColoTensorSpec:
    pg: ProcessGroup
    dist_attr: Optional[_DistSpec] = field(default_factory=ReplicaSpec)
    compute_attr: Optional[ComputeSpec] = None
