"""Sharding policies for the Whisper family.  Parity: reference `colossalai/shardformer/policies/whisper.py:30-560`
(`WhisperModelPolicy`, `WhisperForConditionalGenerationPolicy`, `WhisperForAudioClassificationPolicy`): TP over the
encoder and decoder attention heads / FFN width; the decoder token embedding and the tied output projection are
vocab-parallel."""
from __future__ import annotations

from typing import Dict

import torch.nn as nn

from ..layer.embedding import VocabParallelEmbedding1D
from ..layer.linear import VocabParallelLMHead1D
from .base_policy import ModulePolicyDescription, SubModuleReplacementDescription
from .encdec import EncDecPolicy

__all__ = ["WhisperModelPolicy", "WhisperForConditionalGenerationPolicy", "WhisperForAudioClassificationPolicy"]


class WhisperModelPolicy(EncDecPolicy):
    head_fields = ["encoder_attention_heads", "decoder_attention_heads"]

    def module_policy(self) -> Dict:
        from ...models.whisper import WhisperDecoder, WhisperForConditionalGeneration

        policy = self.block_policies()
        if self.tp == 1:
            return policy
        sc = self.shard_config
        policy[WhisperDecoder] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription(
                "embed_tokens", VocabParallelEmbedding1D,
                kwargs=dict(make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                            fp8_communication=sc.fp8_communication))])
        policy[WhisperForConditionalGeneration] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription(
                "proj_out", VocabParallelLMHead1D,
                kwargs=dict(gather_output=not sc.parallel_output,
                            make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                            fp8_communication=sc.fp8_communication))])
        return policy

    def postprocess(self) -> nn.Module:
        model = super().postprocess()
        head = getattr(model, "proj_out", None)
        if head is not None:
            emb = model.model.decoder.embed_tokens
            if getattr(head, "weight", None) is not None and head.weight.shape == emb.weight.shape:
                head.weight = emb.weight
        return model


class WhisperForConditionalGenerationPolicy(WhisperModelPolicy):
    pass


class WhisperForAudioClassificationPolicy(WhisperModelPolicy):
    head_fields = ["encoder_attention_heads"]
