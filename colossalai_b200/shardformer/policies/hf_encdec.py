"""Policy that shards a user's HuggingFace Whisper encoder-decoder in place (reference `policies/whisper.py:30-300`):
every attention block (encoder self-attention, decoder self- and cross-attention) gets column-parallel q / k / v and a
row-parallel `out_proj`, every feed-forward `fc1` / `fc2` the column / row pair; `WhisperAttention` views k / v with
`self.num_heads`, so that attribute (and `embed_dim`) becomes the local value.  The decoder's token embedding and the
tied `proj_out` head are sharded along the vocabulary identically and re-tied; the convolutional front-end and the
position embeddings stay replicated."""
from __future__ import annotations

from typing import Dict, List

import torch.nn as nn

from ..layer import Linear1D_Col, Linear1D_Row, VocabParallelEmbedding1D, VocabParallelLMHead1D
from .base_policy import ModulePolicyDescription, SubModuleReplacementDescription
from .hf_gpt import _HFTiedDecoderPolicy

__all__ = ["HFWhisperPolicy"]


def _attn(prefix: str, fp8: dict) -> List[SubModuleReplacementDescription]:
    return [SubModuleReplacementDescription(f"{prefix}.q_proj", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription(f"{prefix}.k_proj", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription(f"{prefix}.v_proj", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription(f"{prefix}.out_proj", Linear1D_Row, kwargs=dict(fp8))]


class HFWhisperPolicy(_HFTiedDecoderPolicy):
    """`WhisperModel`, `WhisperForConditionalGeneration`."""

    def config_sanity_check(self) -> None:
        cfg, tp = self.model.config, self.shard_config.tensor_parallel_size
        if self.shard_config.enable_tensor_parallelism:
            assert cfg.encoder_attention_heads % tp == 0 and cfg.decoder_attention_heads % tp == 0, \
                "encoder / decoder attention heads must be divisible by the TP size"
        assert not self.shard_config.enable_sequence_parallelism, \
            "sequence parallelism of HF modules is not supported; build the model from the native zoo (models.hf_io_encdec)"

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        assert cfg.encoder_attention_heads == cfg.decoder_attention_heads, \
            "attribute replacement is per class: encoder and decoder must have the same number of heads"
        fp8 = dict(fp8_communication=sc.fp8_communication)
        ffn = [SubModuleReplacementDescription("fc1", Linear1D_Col, kwargs=dict(fp8)),
               SubModuleReplacementDescription("fc2", Linear1D_Row, kwargs=dict(fp8))]
        policy["WhisperAttention"] = ModulePolicyDescription(attribute_replacement={
            "embed_dim": cfg.d_model // tp, "num_heads": cfg.decoder_attention_heads // tp})
        policy["WhisperEncoderLayer"] = ModulePolicyDescription(sub_module_replacement=_attn("self_attn", fp8) + ffn)
        policy["WhisperDecoderLayer"] = ModulePolicyDescription(
            sub_module_replacement=_attn("self_attn", fp8) + _attn("encoder_attn", fp8) + ffn)
        policy["WhisperDecoder"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("embed_tokens", VocabParallelEmbedding1D, kwargs=self._vocab_kwargs())])
        policy["WhisperForConditionalGeneration"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("proj_out", VocabParallelLMHead1D,
                                            kwargs=dict(gather_output=True, **self._vocab_kwargs()))])
        return policy
