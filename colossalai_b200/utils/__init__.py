from .common import (
    conditional_context,
    disposable,
    ensure_path_exists,
    free_storage,
    get_current_device,
    is_ddp_ignored,
    set_seed,
    get_non_persistent_buffers_set,
)
from .timer import MultiTimer, Timer
from .memory import colo_device_memory_capacity, colo_get_cpu_memory_capacity, colo_set_process_memory_fraction
from .multi_tensor import multi_tensor_applier, MultiTensorApply
from .tensor_detector import TensorDetector
from .rank_recorder import recorder as rank_recorder

__all__ = [
    "conditional_context", "disposable", "ensure_path_exists", "free_storage", "get_current_device",
    "is_ddp_ignored", "set_seed", "get_non_persistent_buffers_set", "MultiTimer", "Timer",
    "colo_device_memory_capacity", "colo_get_cpu_memory_capacity", "colo_set_process_memory_fraction",
    "multi_tensor_applier", "MultiTensorApply", "TensorDetector", "rank_recorder",
]
