from ._operation import (
    AllGather,
    AllToAll,
    AllToAllUneven,
    DPGradScalerIn,
    DPGradScalerOut,
    EPGradScalerIn,
    EPGradScalerOut,
    HierarchicalAllToAll,
    MoeCombine,
    MoeDispatch,
    ReduceScatter,
    all_to_all_uneven,
    moe_cumsum,
)
from .grouped_gemm import grouped_linear

__all__ = ["AllGather", "AllToAll", "AllToAllUneven", "DPGradScalerIn", "DPGradScalerOut", "EPGradScalerIn",
           "EPGradScalerOut", "HierarchicalAllToAll", "MoeCombine", "MoeDispatch", "ReduceScatter",
           "all_to_all_uneven", "moe_cumsum", "grouped_linear"]
