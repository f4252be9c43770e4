from .api import (
    customized_distributed_tensor_to_param,
    distribute_tensor,
    distribute_tensor_with_customization,
    distribute_tensor_with_spec,
    get_device_mesh,
    get_global_shape,
    get_layout,
    get_sharding_spec,
    init_as_dtensor,
    is_customized_distributed_tensor,
    is_distributed_tensor,
    is_shape_consistent,
    is_sharded,
    mark_customized,
    mark_sharded,
    redistribute,
    shard_along,
    shard_colwise,
    shard_rowwise,
    sharded_tensor_to_existing_param,
    sharded_tensor_to_param,
    to_global,
    to_global_for_customized_distributed_tensor,
)
from .comm_spec import CollectiveCommPattern, CommSpec
from .layout import Layout
from .layout_converter import LayoutConverter
from .sharding_spec import DimSpec, ShardingSpec

__all__ = [
    "customized_distributed_tensor_to_param", "distribute_tensor", "distribute_tensor_with_customization",
    "distribute_tensor_with_spec", "get_device_mesh", "get_global_shape", "get_layout", "get_sharding_spec",
    "init_as_dtensor", "is_customized_distributed_tensor", "is_distributed_tensor", "is_shape_consistent",
    "is_sharded", "mark_customized", "mark_sharded", "redistribute", "shard_along", "shard_colwise",
    "shard_rowwise", "sharded_tensor_to_existing_param", "sharded_tensor_to_param", "to_global",
    "to_global_for_customized_distributed_tensor", "CollectiveCommPattern", "CommSpec", "Layout",
    "LayoutConverter", "DimSpec", "ShardingSpec",
]
