"""Shared sharding policy of the encoder / encoder-decoder / vision families built from `models.encdec` blocks
(ViT, T5, Whisper, BLIP-2's vision tower and Q-Former) — tensor parallelism over heads and FFN width:

* self-attention `qkv_proj` -> fused column-parallel ([q|k|v] blocks sharded consistently), `o_proj` -> row-parallel;
* cross-attention `q_proj` -> column-parallel, `kv_proj` -> fused column-parallel ([k|v]), `o_proj` -> row-parallel;
* FFN `up_proj` / `gate_up_proj` -> column-parallel, `down_proj` -> row-parallel;
* `num_heads` divided by the TP size (the blocks derive shapes from the local projection width).

Like the reference's policies for these families (`policies/{vit,t5,whisper,blip2,sam}.py`: "doesn't support
sequence parallelism now, will ignore the sequence parallelism flag") sequence parallelism is switched off.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Type, Union

import torch.nn as nn

from ..layer.linear import Linear1D_Col, Linear1D_Row
from ..layer.qkv_fused_linear import FusedLinear1D_Col
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription

__all__ = ["EncDecPolicy"]


class EncDecPolicy(Policy):
    #: attribute names holding the head counts this family must keep divisible by the TP size
    head_fields: List[str] = []

    def config_sanity_check(self) -> None:
        sc = self.shard_config
        if sc.enable_sequence_parallelism:
            sc.enable_sequence_parallelism = False
            sc.sequence_parallelism_mode = None
            warnings.warn(f"{type(self).__name__}: sequence parallelism is not supported for this family; flag ignored")
        tp = sc.tensor_parallel_size if sc.enable_tensor_parallelism else 1
        cfg = self.model.cfg
        for f in self.head_fields:
            n = getattr(cfg, f)
            assert n % tp == 0, f"{f}={n} must be divisible by the tensor parallel size {tp}"

    def preprocess(self) -> nn.Module:
        return self.model

    @property
    def tp(self) -> int:
        sc = self.shard_config
        return sc.tensor_parallel_size if sc.enable_tensor_parallelism else 1

    def block_policies(self) -> Dict[Union[str, Type[nn.Module]], ModulePolicyDescription]:
        from ...models.encdec import FeedForward, MultiHeadAttention

        if self.tp == 1:
            return {}
        sc = self.shard_config
        common = dict(fp8_communication=sc.fp8_communication)
        tp = self.tp

        def shard_attention(attn: nn.Module) -> None:
            attn.num_heads = attn.num_heads // tp

        policy: Dict = {}
        policy[MultiHeadAttention] = ModulePolicyDescription(
            param_replacement=[shard_attention],
            sub_module_replacement=[
                SubModuleReplacementDescription("qkv_proj", FusedLinear1D_Col, kwargs=dict(num_splits=3, **common),
                                                ignore_if_not_exist=True),
                SubModuleReplacementDescription("q_proj", Linear1D_Col, kwargs=dict(**common),
                                                ignore_if_not_exist=True),
                SubModuleReplacementDescription("kv_proj", FusedLinear1D_Col, kwargs=dict(num_splits=2, **common),
                                                ignore_if_not_exist=True),
                SubModuleReplacementDescription("o_proj", Linear1D_Row, kwargs=dict(**common)),
            ])
        policy[FeedForward] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("gate_up_proj", FusedLinear1D_Col, kwargs=dict(num_splits=2, **common),
                                            ignore_if_not_exist=True),
            SubModuleReplacementDescription("up_proj", Linear1D_Col, kwargs=dict(**common), ignore_if_not_exist=True),
            SubModuleReplacementDescription("down_proj", Linear1D_Row, kwargs=dict(**common)),
        ])
        return policy

    def module_policy(self) -> Dict[Union[str, Type[nn.Module]], ModulePolicyDescription]:
        return self.block_policies()

    def postprocess(self) -> nn.Module:
        sc = self.shard_config
        for m in self.model.modules():
            if hasattr(m, "shard_config"):
                m.shard_config = sc
        return self.model
