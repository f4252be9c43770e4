"""Activation-checkpoint configs.  Parity: reference `colossalai/shardformer/shard/grad_ckpt_config.py:6-80`."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

__all__ = ["GradientCheckpointConfig", "PipelineGradientCheckpointConfig"]


@dataclass
class GradientCheckpointConfig:
    gradient_checkpointing_ratio: float = 0.0

    def get_num_ckpt_layers(self, num_layers: int) -> int:
        return int(self.gradient_checkpointing_ratio * num_layers)


@dataclass
class PipelineGradientCheckpointConfig(GradientCheckpointConfig):
    """Either a global ratio or explicit per-(stage x chunk) layer counts, e.g. [19, 19, 19, 13]."""

    gradient_checkpointing_ratio: Optional[float] = None
    num_ckpt_layers_per_stage: Optional[List[int]] = None

    def __post_init__(self) -> None:
        if self._enable_gradient_checkpointing_ratio:
            if not (0 <= self.gradient_checkpointing_ratio <= 1):
                raise ValueError("gradient_checkpointing_ratio should be in [0, 1]")
        if self._enable_customized_ckpt_layers_per_stage:
            assert all(c >= 0 for c in self.num_ckpt_layers_per_stage)

    @property
    def _enable_gradient_checkpointing_ratio(self) -> bool:
        return self.gradient_checkpointing_ratio is not None

    @property
    def _enable_customized_ckpt_layers_per_stage(self) -> bool:
        return self.num_ckpt_layers_per_stage is not None

    def get_num_ckpt_layers(self, stage: int, num_stages: int, num_layers: int, model_chunk_id: int = 0,
                            num_model_chunks: int = 1) -> int:
        if not self._enable_gradient_checkpointing_ratio and not self._enable_customized_ckpt_layers_per_stage:
            raise RuntimeError("no checkpointed layers specified")
        if self._enable_customized_ckpt_layers_per_stage:
            assert len(self.num_ckpt_layers_per_stage) == num_stages * num_model_chunks
            n = self.num_ckpt_layers_per_stage[stage + model_chunk_id * num_stages]
            assert n <= num_layers
            return n
        return int(self.gradient_checkpointing_ratio * num_layers)
