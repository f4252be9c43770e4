from .shard import (
    GradientCheckpointConfig,
    ModelSharder,
    PipelineGradientCheckpointConfig,
    ShardConfig,
    ShardFormer,
    shard_model,
)

__all__ = ["ShardConfig", "ShardFormer", "ModelSharder", "shard_model", "GradientCheckpointConfig",
           "PipelineGradientCheckpointConfig"]
