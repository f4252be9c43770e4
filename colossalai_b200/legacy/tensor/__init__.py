"""ColoTensor-era distribution specs (reference `colossalai/legacy/tensor`): `ProcessGroup` (tp x dp layout of the
world), `ReplicaSpec` / `ShardSpec`, `ComputePattern` / `ComputeSpec`, `ColoTensorSpec`, `DistSpecManager` (autograd-
aware conversion between specs) and the `colo_op_impl` operator registry."""
from .compute_spec import ComputePattern, ComputeSpec
from .const import TensorType
from .dist_spec_mgr import DistSpecManager
from .distspec import DistPlacementPattern, ReplicaSpec, ShardSpec
from .op_wrapper import colo_op_impl, get_colo_op_impl
from .process_group import ProcessGroup
from .tensor_spec import ColoTensorSpec

__all__ = ["ComputePattern", "ComputeSpec", "TensorType", "DistSpecManager", "DistPlacementPattern", "ReplicaSpec",
           "ShardSpec", "colo_op_impl", "get_colo_op_impl", "ProcessGroup", "ColoTensorSpec"]
