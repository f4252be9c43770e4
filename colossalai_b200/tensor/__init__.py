from . import d_tensor, moe_tensor, padded_tensor
from .d_tensor import (
    distribute_tensor,
    is_distributed_tensor,
    is_sharded,
    redistribute,
    shard_colwise,
    shard_rowwise,
    to_global,
)
from .colo_tensor import ColoParameter, ColoTensor
from .param_op_hook import ColoParamOpHook, ColoParamOpHookManager
from .padded_tensor import is_padded_tensor, to_padded_tensor, to_unpadded_tensor

__all__ = ["ColoParameter", "ColoTensor", "ColoParamOpHook", "ColoParamOpHookManager", "d_tensor", "moe_tensor", "padded_tensor", "distribute_tensor", "is_distributed_tensor", "is_sharded",
           "redistribute", "shard_colwise", "shard_rowwise", "to_global", "is_padded_tensor", "to_padded_tensor",
           "to_unpadded_tensor"]
