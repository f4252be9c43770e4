"""Per-node analysis record (reference `_analyzer/fx/node_util.py` `MetaInfo`)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Tuple

__all__ = ["MetaInfo"]


@dataclass
class MetaInfo:
    fwd_flop: int = 0
    bwd_flop: int = 0
    output_bytes: int = 0            # activation produced by the node
    saved_bytes: int = 0             # tensors autograd keeps alive for the node's backward
    param_bytes: int = 0
    outputs: Any = None              # (shape, dtype) tree of the result
    is_inplace: bool = False

    @property
    def fwd_time_key(self) -> Tuple[int, int]:
        return self.fwd_flop, self.output_bytes

    def accumulate_activation(self) -> int:
        return self.saved_bytes if self.saved_bytes else self.output_bytes
