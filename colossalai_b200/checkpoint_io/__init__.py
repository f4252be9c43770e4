from .checkpoint_io_base import CheckpointIO
from .general_checkpoint_io import GeneralCheckpointIO
from .index_file import CheckpointIndexFile

__all__ = ["CheckpointIO", "CheckpointIndexFile", "GeneralCheckpointIO", "HybridParallelCheckpointIO",
           "MoECheckpointIO"]


def __getattr__(name):
    if name == "HybridParallelCheckpointIO":
        from .hybrid_parallel_checkpoint_io import HybridParallelCheckpointIO

        return HybridParallelCheckpointIO
    if name == "MoECheckpointIO":
        from .moe_checkpoint import MoECheckpointIO

        return MoECheckpointIO
    raise AttributeError(name)
