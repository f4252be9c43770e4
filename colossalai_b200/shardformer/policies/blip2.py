"""Sharding policies for BLIP-2.  Parity: reference `colossalai/shardformer/policies/blip2.py:20-420`
(`Blip2ModelPolicy`, `Blip2ForConditionalGenerationPolicy`): the vision tower and the Q-Former are sharded with the
encoder block policy, the language model with the decoder-stack policy (vocab-parallel embedding + LM head)."""
from __future__ import annotations

from typing import Dict

import torch.nn as nn

from .encdec import EncDecPolicy
from .transformer import TransformerPolicy

__all__ = ["Blip2ModelPolicy", "Blip2ForConditionalGenerationPolicy"]


class Blip2ModelPolicy(EncDecPolicy):
    head_fields = ["vision_heads", "qformer_heads"]

    def _lm_policy(self) -> TransformerPolicy:
        p = TransformerPolicy()
        p.set_model(self.model.language_model)
        p.shard_config = self.shard_config
        return p

    def config_sanity_check(self) -> None:
        super().config_sanity_check()
        self._lm_policy().config_sanity_check()

    def module_policy(self) -> Dict:
        policy = self.block_policies()
        policy.update(self._lm_policy().module_policy())
        return policy

    def postprocess(self) -> nn.Module:
        model = super().postprocess()
        self._lm_policy().postprocess()
        return model


class Blip2ForConditionalGenerationPolicy(Blip2ModelPolicy):
    pass
