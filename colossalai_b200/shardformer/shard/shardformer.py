"""ShardFormer façade.  Parity: reference `colossalai/shardformer/shard/shardformer.py:14-56`."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch.nn as nn
from torch import Tensor

from ..policies.base_policy import Policy
from .shard_config import ShardConfig
from .sharder import ModelSharder

__all__ = ["ShardFormer"]


class ShardFormer:
    """`ShardFormer(shard_config).optimize(model, policy=None) -> (model, shared_params)`"""

    def __init__(self, shard_config: ShardConfig) -> None:
        self.shard_config = shard_config

    def optimize(self, model: nn.Module, policy: Policy = None) -> Tuple[nn.Module, List[Dict[int, Tensor]]]:
        sharder = ModelSharder(model=model, shard_config=self.shard_config, policy=policy)
        shared_params = sharder.shard()
        return model, shared_params
