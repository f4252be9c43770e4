"""Policies that shard a user's HuggingFace BERT / ViT encoder in place (reference `policies/bert.py:30-260`): query / key /
value and `intermediate.dense` become column-parallel, `attention.output.dense` and `output.dense` row-parallel, the
word embedding vocab-parallel; the attention module sizes its head views with `-1`, so only the bookkeeping attributes
(`num_attention_heads`, `all_head_size`) are replaced.  Covers `BertModel` and the heads that sit on the pooled /
sequence output (sequence / token classification, question answering, multiple choice, next-sentence prediction).  The
masked-LM heads tie a biased decoder to the word embedding and are not handled here (native zoo: the `bert` row of `_family_table.py`)."""
from __future__ import annotations

from typing import Dict, List

import torch.nn as nn

from ..layer import Linear1D_Col, Linear1D_Row, VocabParallelEmbedding1D
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription

__all__ = ["HFBertPolicy", "HFViTPolicy"]


class HFBertPolicy(Policy):
    def config_sanity_check(self) -> None:
        cfg = self.model.config
        if self.shard_config.enable_tensor_parallelism:
            assert cfg.num_attention_heads % self.shard_config.tensor_parallel_size == 0, \
                "num_attention_heads must be divisible by the TP size"
        assert not self.shard_config.enable_sequence_parallelism, \
            "sequence parallelism of HF modules is not supported; build the model from the native zoo (models.hf_io)"

    def preprocess(self) -> nn.Module:
        return self.model

    def postprocess(self) -> nn.Module:
        return self.model

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        fp8 = dict(fp8_communication=sc.fp8_communication)
        policy["BertSelfAttention"] = ModulePolicyDescription(attribute_replacement={
            "num_attention_heads": cfg.num_attention_heads // tp, "all_head_size": cfg.hidden_size // tp})
        policy["BertLayer"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("attention.self.query", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("attention.self.key", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("attention.self.value", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("attention.output.dense", Linear1D_Row, kwargs=dict(fp8)),
            SubModuleReplacementDescription("intermediate.dense", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("output.dense", Linear1D_Row, kwargs=dict(fp8)),
        ])
        policy["BertEmbeddings"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription(
                "word_embeddings", VocabParallelEmbedding1D,
                kwargs=dict(make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                            fp8_communication=sc.fp8_communication))])
        return policy

    def get_held_layers(self) -> List[nn.Module]:
        if self.pipeline_stage_manager is not None:
            raise NotImplementedError("pipeline parallelism of HuggingFace modules: import the weights into the native "
                                      "zoo (`models.hf_io.load_hf_checkpoint`) and use its policy")
        return []

    def get_shared_params(self):
        return []


class HFViTPolicy(HFBertPolicy):
    """`ViTModel`, `ViTForImageClassification` (reference `policies/vit.py`): the same column / row pattern on
    `attention.attention.{query,key,value}`, `attention.output.dense`, `intermediate.dense`, `output.dense`;
    `ViTSelfAttention` reshapes with `num_attention_heads` and `all_head_size`, so both become local values.  Patch and
    position embeddings stay replicated."""

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        fp8 = dict(fp8_communication=sc.fp8_communication)
        policy["ViTSelfAttention"] = ModulePolicyDescription(attribute_replacement={
            "num_attention_heads": cfg.num_attention_heads // tp, "all_head_size": cfg.hidden_size // tp})
        policy["ViTLayer"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("attention.attention.query", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("attention.attention.key", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("attention.attention.value", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("attention.output.dense", Linear1D_Row, kwargs=dict(fp8)),
            SubModuleReplacementDescription("intermediate.dense", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("output.dense", Linear1D_Row, kwargs=dict(fp8)),
        ])
        return policy
