from .flop_tensor import flop_count, flop_mapping
from .meta_tensor import MetaTensor, MetaTensorMode

__all__ = ["MetaTensor", "MetaTensorMode", "flop_count", "flop_mapping"]
