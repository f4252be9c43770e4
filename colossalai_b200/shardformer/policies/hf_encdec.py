"""Policies that shard a user's HuggingFace encoder-decoder (Whisper, T5) in place.

Whisper (reference `policies/whisper.py:30-300`):
every attention block (encoder self-attention, decoder self- and cross-attention) gets column-parallel q / k / v and a
row-parallel `out_proj`, every feed-forward `fc1` / `fc2` the column / row pair; `WhisperAttention` views k / v with
`self.num_heads`, so that attribute (and `embed_dim`) becomes the local value.  The decoder's token embedding and the
tied `proj_out` head are sharded along the vocabulary identically and re-tied; the convolutional front-end and the
position embeddings stay replicated."""
from __future__ import annotations

from typing import Dict, List

import torch.nn as nn

from ..layer import Embedding1D, Linear1D_Col, Linear1D_Row, VocabParallelEmbedding1D, VocabParallelLMHead1D
from .base_policy import ModulePolicyDescription, SubModuleReplacementDescription
from .hf_gpt import _HFTiedDecoderPolicy

__all__ = ["HFWhisperPolicy", "HFT5Policy"]


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


class HFT5Policy(_HFTiedDecoderPolicy):
    """`T5Model`, `T5ForConditionalGeneration`, `T5EncoderModel` (reference `policies/t5.py:30-360`).

    T5 specifics: (1) the relative position bias is an `nn.Embedding(buckets, n_heads)` owned by the first block of each
    stack and its output is passed on to every later block - it is sharded along the HEAD dimension without gathering
    (`Embedding1D(gather_output=False)`), so every block adds the bias of its local heads; (2) `n_heads` / `inner_dim`
    become local values (used when a zero bias is synthesised); (3) `shared`, `encoder.embed_tokens` and
    `decoder.embed_tokens` are one module: it is replaced once and the three references are re-pointed; the tied LM head
    follows the vocabulary shard (T5 rescales the decoder output itself when the weights are tied)."""

    def config_sanity_check(self) -> None:
        cfg, tp = self.model.config, self.shard_config.tensor_parallel_size
        if self.shard_config.enable_tensor_parallelism:
            assert cfg.num_heads % tp == 0, "num_heads must be divisible by the TP size"
        assert not self.shard_config.enable_sequence_parallelism, \
            "sequence parallelism of HF modules is not supported; build the model from the native zoo (models.hf_io_encdec)"

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        fp8 = dict(fp8_communication=sc.fp8_communication)

        def attn(prefix: str) -> List[SubModuleReplacementDescription]:
            return [SubModuleReplacementDescription(f"{prefix}.q", Linear1D_Col, kwargs=dict(fp8)),
                    SubModuleReplacementDescription(f"{prefix}.k", Linear1D_Col, kwargs=dict(fp8)),
                    SubModuleReplacementDescription(f"{prefix}.v", Linear1D_Col, kwargs=dict(fp8)),
                    SubModuleReplacementDescription(f"{prefix}.o", Linear1D_Row, kwargs=dict(fp8)),
                    SubModuleReplacementDescription(f"{prefix}.relative_attention_bias", Embedding1D,
                                                    kwargs=dict(gather_output=False), ignore_if_not_exist=True)]

        policy["T5Attention"] = ModulePolicyDescription(attribute_replacement={
            "n_heads": cfg.num_heads // tp, "inner_dim": cfg.num_heads * cfg.d_kv // tp})
        policy["T5LayerSelfAttention"] = ModulePolicyDescription(sub_module_replacement=attn("SelfAttention"))
        policy["T5LayerCrossAttention"] = ModulePolicyDescription(sub_module_replacement=attn("EncDecAttention"))
        policy["T5DenseActDense"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("wi", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("wo", Linear1D_Row, kwargs=dict(fp8))])
        policy["T5DenseGatedActDense"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("wi_0", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("wi_1", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("wo", Linear1D_Row, kwargs=dict(fp8))])
        emb = [SubModuleReplacementDescription("shared", VocabParallelEmbedding1D, kwargs=self._vocab_kwargs())]
        policy["T5Model"] = ModulePolicyDescription(sub_module_replacement=emb)
        policy["T5EncoderModel"] = ModulePolicyDescription(sub_module_replacement=emb)
        policy["T5ForConditionalGeneration"] = ModulePolicyDescription(sub_module_replacement=emb + [
            SubModuleReplacementDescription("lm_head", VocabParallelLMHead1D,
                                            kwargs=dict(gather_output=True, **self._vocab_kwargs()))])
        return policy

    def postprocess(self) -> nn.Module:
        if self.shard_config.enable_tensor_parallelism:
            shared = self.model.shared
            for stack in ("encoder", "decoder"):
                mod = getattr(self.model, stack, None)
                if mod is not None and hasattr(mod, "embed_tokens"):
                    mod.embed_tokens = shared
        return super().postprocess()
