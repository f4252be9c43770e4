"""Policies that shard a user's HuggingFace decoder model IN PLACE (transformers >= 4.48 module layout: separate
`q_proj / k_proj / v_proj / o_proj`, `gate_proj / up_proj / down_proj`, `embed_tokens`, `lm_head`).

Our own model zoo is born parallel (`policies/transformer.py`); this file is the reference's other use case - take an
existing `nn.Module` the framework knows nothing about and rewrite it through the generic `ModelSharder` machinery
(sub-module replacement by dotted suffix, method replacement, attribute replacement), exactly like the reference's
`policies/llama.py:30-175` + `shard/sharder.py:77-160`:

  * tensor parallelism: q/k/v/gate/up -> `Linear1D_Col`, o/down -> `Linear1D_Row`, `embed_tokens` ->
    `VocabParallelEmbedding1D`, `lm_head` -> `VocabParallelLMHead1D` (gathered logits, so the model's own loss works);
    HF attention modules size their head views with `-1`, so the sharded projections are all they need;
  * method replacement: every `*RMSNorm.forward` is rebound to our fused RMSNorm kernel (`ops.rms_norm`);
  * data parallel / ZeRO plugins need no policy at all.
Pipeline parallelism and sequence parallelism of HF modules are NOT provided here (use the native zoo + `hf_io` weight
import for those): `get_held_layers` raises if a stage manager is present.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

from ..layer import Linear1D_Col, Linear1D_Row, VocabParallelEmbedding1D, VocabParallelLMHead1D
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription

__all__ = ["HFDecoderPolicy", "HF_FAMILIES"]

# transformers module path -> class-name prefix
HF_FAMILIES = {"llama": "Llama", "mistral": "Mistral", "qwen2": "Qwen2", "qwen3": "Qwen3", "cohere": "Cohere"}


def _fused_rmsnorm_forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
    from ... import ops

    eps = getattr(self, "variance_epsilon", getattr(self, "eps", 1e-6))
    return ops.rms_norm(hidden_states, self.weight, eps)


class HFDecoderPolicy(Policy):
    """Works for `<Family>Model`, `<Family>ForCausalLM` of the families in HF_FAMILIES (matched by class NAME, so the
    policy never imports transformers itself)."""

    def config_sanity_check(self) -> None:
        cfg = self.model.config
        tp = self.shard_config.tensor_parallel_size
        if self.shard_config.enable_tensor_parallelism:
            assert cfg.num_attention_heads % tp == 0, "num_attention_heads must be divisible by the TP size"
            kv = getattr(cfg, "num_key_value_heads", cfg.num_attention_heads)
            assert kv % tp == 0, "num_key_value_heads must be divisible by the TP size"
        assert not self.shard_config.enable_sequence_parallelism, \
            "sequence parallelism of HF modules is not supported; build the model from the native zoo (models.hf_io)"

    def preprocess(self) -> nn.Module:
        self.tie_weight = self.tie_weight_check()
        return self.model

    def _prefix(self) -> str:
        name = self.model.__class__.__name__
        for p in HF_FAMILIES.values():
            if name.startswith(p):
                return p
        raise NotImplementedError(f"{name} is not a supported HuggingFace decoder family {sorted(HF_FAMILIES.values())}")

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        fam = self._prefix()
        policy: Dict[str, ModulePolicyDescription] = {}
        if sc.enable_tensor_parallelism:
            col = dict(fp8_communication=sc.fp8_communication)
            policy[f"{fam}DecoderLayer"] = ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription("self_attn.q_proj", Linear1D_Col, kwargs=dict(col)),
                SubModuleReplacementDescription("self_attn.k_proj", Linear1D_Col, kwargs=dict(col)),
                SubModuleReplacementDescription("self_attn.v_proj", Linear1D_Col, kwargs=dict(col)),
                SubModuleReplacementDescription("self_attn.o_proj", Linear1D_Row, kwargs=dict(col)),
                SubModuleReplacementDescription("mlp.gate_proj", Linear1D_Col, kwargs=dict(col)),
                SubModuleReplacementDescription("mlp.up_proj", Linear1D_Col, kwargs=dict(col)),
                SubModuleReplacementDescription("mlp.down_proj", Linear1D_Row, kwargs=dict(col)),
            ])
            policy[f"{fam}Model"] = ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription(
                    "embed_tokens", VocabParallelEmbedding1D,
                    kwargs=dict(make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                                fp8_communication=sc.fp8_communication))])
            policy[f"{fam}ForCausalLM"] = ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription(
                    "lm_head", VocabParallelLMHead1D,
                    kwargs=dict(gather_output=True, make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                                fp8_communication=sc.fp8_communication))])
        if sc.enable_fused_normalization:
            policy[f"{fam}RMSNorm"] = ModulePolicyDescription(method_replacement={"forward": _fused_rmsnorm_forward})
        return policy

    def postprocess(self) -> nn.Module:
        if getattr(self, "tie_weight", False) and self.shard_config.enable_tensor_parallelism:
            # both sides are sharded along the vocab dimension identically: re-tie the local shards
            emb = self.model.get_input_embeddings()
            head = self.model.get_output_embeddings()
            if head is not None and emb is not None and head.weight.shape == emb.weight.shape:
                head.weight = emb.weight
        return self.model

    def get_held_layers(self) -> List[nn.Module]:
        if self.pipeline_stage_manager is not None:
            raise NotImplementedError("pipeline parallelism of HuggingFace modules: import the weights into the native "
                                      "zoo (`models.hf_io.load_hf_checkpoint`) and use its policy")
        return []

    def get_shared_params(self):
        return []
