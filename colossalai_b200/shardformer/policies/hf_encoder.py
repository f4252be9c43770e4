"""Policies that shard a user's HuggingFace BERT / ViT encoder in place (reference `policies/bert.py:30-260`): query / key /
value and `intermediate.dense` become column-parallel, `attention.output.dense` and `output.dense` row-parallel, the
word embedding vocab-parallel; the attention module sizes its head views with `-1`, so only the bookkeeping attributes
(`num_attention_heads`, `all_head_size`) are replaced.  Covers `BertModel` and the heads that sit on the pooled /
sequence output (sequence / token classification, question answering, multiple choice, next-sentence prediction).
Pipeline stages (1F1B): a stage keeps its slice of `encoder.layer`; behind the first stage the `embeddings` module is a
pass-through of `inputs_embeds` (it would add positions / token types and normalise again), the pooler exists on the
last stage only, and the LAST stage simply calls the model's own forward with `inputs_embeds=<received hidden states>` -
pooler, task head and loss are the user's code, whatever the head is.  The
masked-LM heads tie a biased decoder to the word embedding and are not handled here (native zoo: the `bert` row of `_family_table.py`)."""
from __future__ import annotations

from typing import Dict, List

import torch.nn as nn

from .._utils import getattr_, setattr_
from ..layer import Linear1D_Col, Linear1D_Row, VocabParallelEmbedding1D
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription

__all__ = ["HFBertPolicy", "HFViTPolicy"]


class _PassThroughEmbeddings(nn.Module):
    def forward(self, input_ids=None, inputs_embeds=None, **kwargs):
        return inputs_embeds


def _encoder_stage_forward(self, input_ids=None, hidden_states=None, attention_mask=None, labels=None, **kwargs):
    """Stage-aware forward bound to the user's BERT model under pipeline parallelism."""
    sm = self._cb200_stage_manager
    inputs = dict(input_ids=input_ids) if sm.is_first_stage() else dict(inputs_embeds=hidden_states)
    if sm.is_last_stage():
        out = self._cb200_forward(**inputs, attention_mask=attention_mask, labels=labels, **kwargs)
        return {k: v for k, v in out.items() if v is not None}
    backbone = self._cb200_backbone
    extra = {k: v for k, v in kwargs.items() if k in ("token_type_ids", "position_ids")} if sm.is_first_stage() else {}
    out = type(backbone).forward(backbone, **inputs, attention_mask=attention_mask, **extra)
    return {"hidden_states": out.last_hidden_state}


class HFBertPolicy(Policy):
    _pp_layers = "encoder.layer"

    def _backbone(self) -> nn.Module:
        return getattr(self.model, "bert", self.model)

    def _install_pipeline_stage(self) -> None:
        sm = self.pipeline_stage_manager
        if sm is None or sm.num_stages == 1:
            return
        from types import MethodType

        backbone = self._backbone()
        setattr_(backbone, self._pp_layers, nn.ModuleList(self._held_layers))
        if not sm.is_first_stage():
            backbone.embeddings = _PassThroughEmbeddings()
        if not sm.is_last_stage():
            backbone.pooler = None
        self.model._cb200_forward = self.model.forward            # the user's forward (bound), used by the last stage
        self.model._cb200_backbone = backbone
        self.model._cb200_stage_manager = sm
        self.model.forward = MethodType(_encoder_stage_forward, self.model)

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
        self._install_pipeline_stage()
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
        sm = self.pipeline_stage_manager
        if sm is None:
            return []
        assert not sm.is_interleave, "HF modules support the 1F1B schedule (one model chunk per stage)"
        backbone = self._backbone()
        layers = getattr_(backbone, self._pp_layers)
        start, end = sm.get_stage_index(sm.distribute_layers(len(layers)))
        self._held_layers = list(layers[start:end])
        held: List[nn.Module] = list(self._held_layers)
        if sm.is_first_stage():
            held.append(backbone.embeddings)
        if sm.is_last_stage():
            # the pooler and everything outside the backbone (dropout, classifier, qa_outputs, ...)
            if getattr(backbone, "pooler", None) is not None:
                held.append(backbone.pooler)
            held += [m for n, m in self.model.named_children() if m is not backbone]
        return held

    def get_shared_params(self):
        return []


class HFViTPolicy(HFBertPolicy):
    """`ViTModel`, `ViTForImageClassification` (reference `policies/vit.py`): the same column / row pattern on
    `attention.attention.{query,key,value}`, `attention.output.dense`, `intermediate.dense`, `output.dense`;
    `ViTSelfAttention` reshapes with `num_attention_heads` and `all_head_size`, so both become local values.  Patch and
    position embeddings stay replicated.  `ViTModel` takes pixels only (no `inputs_embeds`): pipeline stages of ViT are
    left to the native zoo."""

    def postprocess(self) -> nn.Module:
        return self.model

    def get_held_layers(self) -> List[nn.Module]:
        if self.pipeline_stage_manager is not None:
            raise NotImplementedError("pipeline parallelism of HuggingFace ViT: import the weights into the native zoo "
                                      "(`models.hf_io`) and use its policy")
        return []

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
