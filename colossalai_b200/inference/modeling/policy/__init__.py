"""Inference-time sharding policies + the `model_policy_map` the engine consults.

Parity: reference `colossalai/inference/modeling/policy/__init__.py` (`model_policy_map` with `nopadding_llama`,
`nopadding_baichuan`, `glide_llama`, `pixart_alpha`, `stable_diffusion_3`) and the policy files next to it."""
from __future__ import annotations

from ....shardformer.policies.baichuan import BaichuanForCausalLMPolicy
from ....shardformer.policies.transformer import TransformerForCausalLMPolicy

__all__ = ["NoPaddingLlamaModelInferPolicy", "NoPaddingBaichuanModelInferPolicy", "GlideLlamaModelPolicy",
           "PixArtAlphaInferPolicy", "StableDiffusion3InferPolicy", "model_policy_map", "get_infer_policy"]


class NoPaddingLlamaModelInferPolicy(TransformerForCausalLMPolicy):
    """TP policy of the un-padded Llama-like inference models (`parallel_output` off: the sampler needs full logits)."""

    def config_sanity_check(self) -> None:
        self.shard_config.parallel_output = False
        super().config_sanity_check()


class NoPaddingBaichuanModelInferPolicy(BaichuanForCausalLMPolicy):
    def config_sanity_check(self) -> None:
        self.shard_config.parallel_output = False
        super().config_sanity_check()


class GlideLlamaModelPolicy(NoPaddingLlamaModelInferPolicy):
    """The GLIDE drafter: same sharding as Llama; the cross-attention projections stay replicated (they index the
    LARGE model's head space, which the drafter reads whole)."""


class _DiTPolicy:
    """Diffusion transformers are sharded by PATCH (Distrifusion) rather than by weight: nothing to replace, the
    engine wraps the DiT's layers with `layers.distrifusion` modules instead."""

    patch_parallel = True


class PixArtAlphaInferPolicy(_DiTPolicy):
    pass


class StableDiffusion3InferPolicy(_DiTPolicy):
    pass


model_policy_map = {
    "nopadding_llama": NoPaddingLlamaModelInferPolicy,
    "nopadding_baichuan": NoPaddingBaichuanModelInferPolicy,
    "glide_llama": GlideLlamaModelPolicy,
    "pixart_alpha": PixArtAlphaInferPolicy,
    "PixArtAlphaPipeline": PixArtAlphaInferPolicy,
    "stable_diffusion_3": StableDiffusion3InferPolicy,
    "StableDiffusion3Pipeline": StableDiffusion3InferPolicy,
}


def get_infer_policy(model_type: str):
    """`llama`/`mistral`/`qwen2`... -> no-padding llama policy, `baichuan` -> its own."""
    if model_type in model_policy_map:
        return model_policy_map[model_type]
    return model_policy_map["nopadding_baichuan" if model_type == "baichuan" else "nopadding_llama"]
