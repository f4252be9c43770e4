"""How an operator should treat a parameter (reference `legacy/tensor/compute_spec.py`)."""
from enum import Enum

__all__ = ["ComputePattern", "ComputeSpec"]


class ComputePattern(Enum):
    TP1D = 0
    TP2D = 1
    TP2P5D = 2
    TP3D = 3


class ComputeSpec:
    """`output_replicate`: gather the (column-parallel) result back to a replicated tensor after the op."""

    def __init__(self, compute_pattern: ComputePattern) -> None:
        assert isinstance(compute_pattern, ComputePattern)
        self.compute_pattern = compute_pattern
        self.output_replicate = True

    def set_output_replicate(self, flag: bool = True) -> None:
        self.output_replicate = flag

    def __repr__(self) -> str:
        return f"ComputeSpec(pattern={self.compute_pattern.name}, replicate_output={self.output_replicate})"
