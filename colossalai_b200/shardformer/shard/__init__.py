from .grad_ckpt_config import GradientCheckpointConfig, PipelineGradientCheckpointConfig
from .shard_config import ShardConfig
from .sharder import ModelSharder, shard_model
from .shardformer import ShardFormer

__all__ = ["ShardConfig", "ModelSharder", "shard_model", "ShardFormer", "GradientCheckpointConfig",
           "PipelineGradientCheckpointConfig"]
