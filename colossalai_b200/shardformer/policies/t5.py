"""Sharding policies for the T5 family.  Parity: reference `colossalai/shardformer/policies/t5.py:34-560`
(`T5ModelPolicy`, `T5ForConditionalGenerationPolicy`, `T5EncoderPolicy`, `T5ForTokenClassificationPolicy`).

On top of the block policy: the shared token embedding is vocab-parallel, the LM head is a vocab-parallel column
linear (logits stay vocab-sharded for the distributed cross entropy when `parallel_output`), and each stack's
relative-position table `[buckets, heads]` is sharded along the heads so every rank materialises only the bias of its
own heads."""
from __future__ import annotations

from typing import Dict

import torch.nn as nn

from ..layer.embedding import Embedding1D, VocabParallelEmbedding1D
from ..layer.linear import VocabParallelLMHead1D
from .base_policy import ModulePolicyDescription, SubModuleReplacementDescription
from .encdec import EncDecPolicy

__all__ = ["T5ModelPolicy", "T5ForConditionalGenerationPolicy", "T5EncoderModelPolicy", "T5EncoderPolicy",
           "T5ForTokenClassificationPolicy"]


class T5ModelPolicy(EncDecPolicy):
    head_fields = ["num_heads"]

    def module_policy(self) -> Dict:
        from ...models.t5 import T5Stack, _T5Base

        policy = self.block_policies()
        if self.tp == 1:
            return policy
        sc = self.shard_config
        policy[T5Stack] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("relative_attention_bias", Embedding1D,
                                            kwargs=dict(gather_output=False))])
        subs = [SubModuleReplacementDescription(
            "shared", VocabParallelEmbedding1D,
            kwargs=dict(make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                        fp8_communication=sc.fp8_communication))]
        if hasattr(self.model, "lm_head"):
            subs.append(SubModuleReplacementDescription(
                "lm_head", VocabParallelLMHead1D,
                kwargs=dict(gather_output=not sc.parallel_output,
                            make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                            fp8_communication=sc.fp8_communication)))
        policy[_T5Base] = ModulePolicyDescription(sub_module_replacement=subs)
        return policy

    def postprocess(self) -> nn.Module:
        model = super().postprocess()
        cfg = model.cfg
        head = getattr(model, "lm_head", None)
        if cfg.tie_word_embeddings and head is not None and getattr(head, "weight", None) is not None \
                and head.weight.shape == model.shared.weight.shape:
            head.weight = model.shared.weight
        return model


class T5ForConditionalGenerationPolicy(T5ModelPolicy):
    pass


class T5EncoderModelPolicy(T5ModelPolicy):
    pass


T5EncoderPolicy = T5EncoderModelPolicy


class T5ForTokenClassificationPolicy(T5ModelPolicy):
    pass
