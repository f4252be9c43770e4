"""Sharding policy for Segment Anything.  Parity: reference `colossalai/shardformer/policies/sam.py:14-260`
(`SamModelPolicy`): column/row parallel q/k/v/out projections and MLPs of the vision encoder, and the q/k/v/out
projections + MLP of the mask decoder's two-way transformer; the replicated decomposed relative-position tables get
their partial gradients summed over the TP group inside the attention forward."""
from __future__ import annotations

from typing import Dict

import torch.nn as nn

from ..layer.linear import Linear1D_Col, Linear1D_Row
from ..layer.qkv_fused_linear import FusedLinear1D_Col
from .base_policy import ModulePolicyDescription, SubModuleReplacementDescription
from .encdec import EncDecPolicy

__all__ = ["SamModelPolicy"]


class SamModelPolicy(EncDecPolicy):
    head_fields = ["vision_heads", "decoder_heads"]

    def module_policy(self) -> Dict:
        from ...models.sam import SamAttention, SamMLP, SamVisionAttention

        if self.tp == 1:
            return {}
        tp = self.tp
        common = dict(fp8_communication=self.shard_config.fp8_communication)

        def shard_heads(attn: nn.Module) -> None:
            attn.num_heads = attn.num_heads // tp

        return {
            SamVisionAttention: ModulePolicyDescription(param_replacement=[shard_heads], sub_module_replacement=[
                SubModuleReplacementDescription("qkv_proj", FusedLinear1D_Col, kwargs=dict(num_splits=3, **common)),
                SubModuleReplacementDescription("o_proj", Linear1D_Row, kwargs=dict(**common))]),
            SamAttention: ModulePolicyDescription(param_replacement=[shard_heads], sub_module_replacement=[
                SubModuleReplacementDescription("q_proj", Linear1D_Col, kwargs=dict(**common)),
                SubModuleReplacementDescription("k_proj", Linear1D_Col, kwargs=dict(**common)),
                SubModuleReplacementDescription("v_proj", Linear1D_Col, kwargs=dict(**common)),
                SubModuleReplacementDescription("o_proj", Linear1D_Row, kwargs=dict(**common))]),
            SamMLP: ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription("up_proj", Linear1D_Col, kwargs=dict(**common)),
                SubModuleReplacementDescription("down_proj", Linear1D_Row, kwargs=dict(**common))]),
        }
