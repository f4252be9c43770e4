"""The sharding policy of our generic transformer (every decoder/encoder family in `colossalai_b200.models`).

Parity: reference `colossalai/shardformer/policies/llama.py:30-400` (and the 20 sibling policies): which linears are
column / row parallel, head-count attribute replacement, vocab-parallel embedding + LM head, fused norms with
`sp_partial_derived`, pipeline `get_held_layers` / `get_shared_params`.  Because the model's forward is already
parallel-aware, the policy needs no method replacement — only sub-module and attribute replacement.
"""
from __future__ import annotations

from typing import Dict, List, Type, Union

import torch.nn as nn
from torch import Tensor

from ...parallel import comm
from ..layer.embedding import PaddingEmbedding, VocabParallelEmbedding1D
from ..layer.linear import Linear1D_Col, Linear1D_Row, LinearWithGradAccum, PaddingLMHead, VocabParallelLMHead1D
from ..layer.normalization import FusedLayerNorm, FusedRMSNorm
from ..layer.qkv_fused_linear import FusedLinear1D_Col
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription

__all__ = ["TransformerPolicy", "TransformerForCausalLMPolicy"]


class TransformerPolicy(Policy):
    def config_sanity_check(self) -> None:
        sc = self.shard_config
        cfg = self.model.cfg
        tp = sc.tensor_parallel_size
        if sc.enable_tensor_parallelism and tp > 1:
            assert cfg.num_attention_heads % tp == 0, (
                f"num_attention_heads {cfg.num_attention_heads} must be divisible by tp {tp}")
            assert cfg.num_key_value_heads % tp == 0, (
                f"num_key_value_heads {cfg.num_key_value_heads} must be divisible by tp {tp} "
                "(KV-head replication is not supported)")
        if sc.sp_mode == "all_to_all":
            sp = sc.sequence_parallel_size
            assert (cfg.num_attention_heads // tp) % sp == 0 and (cfg.num_key_value_heads // tp) % sp == 0, (
                "all_to_all sequence parallelism needs heads/tp divisible by sp")
        if sc.sp_mode == "ring_attn":
            assert cfg.causal, "ring attention supports causal LMs only"

    def preprocess(self) -> nn.Module:
        return self.model

    def module_policy(self) -> Dict[Union[str, Type[nn.Module]], ModulePolicyDescription]:
        from ...models.moe import SparseMoE
        from ...models.transformer import (Attention, DecoderLayer, MLAttention, MLP, TransformerLMHeadModel,
                                           TransformerModel)

        sc = self.shard_config
        cfg = self.model.cfg
        tp = sc.tensor_parallel_size if sc.enable_tensor_parallelism else 1
        sp_mode = sc.sp_mode
        lin_sp = sp_mode if sp_mode in ("split_gather", "ring") else None
        sp_partial = lin_sp is not None
        zbv = sc.use_zbv
        common = dict(seq_parallel_mode=lin_sp, seq_parallel_dim=0, fp8_communication=sc.fp8_communication, use_zbv=zbv)
        policy: Dict = {}
        norm_cls = FusedRMSNorm if cfg.norm_type == "rms" else FusedLayerNorm

        if sc.enable_tensor_parallelism and tp > 1:
            policy[Attention] = ModulePolicyDescription(
                attribute_replacement={"num_heads": cfg.num_attention_heads // tp,
                                       "num_kv_heads": cfg.num_key_value_heads // tp},
                sub_module_replacement=[
                    SubModuleReplacementDescription(
                        "qkv_proj", FusedLinear1D_Col,
                        kwargs=dict(split_sizes=[cfg.q_size, cfg.kv_size, cfg.kv_size], **common)),
                    SubModuleReplacementDescription("o_proj", Linear1D_Row, kwargs=dict(**common)),
                ])
            if cfg.use_mla:
                assert lin_sp is None, "multi-head latent attention: sequence parallelism is not supported yet"
                policy[MLAttention] = ModulePolicyDescription(
                    attribute_replacement={"num_heads": cfg.num_attention_heads // tp,
                                           "num_kv_heads": cfg.num_attention_heads // tp},
                    sub_module_replacement=[
                        SubModuleReplacementDescription("q_b_proj" if cfg.q_lora_rank is not None else "q_proj",
                                                        Linear1D_Col, kwargs=dict(**common)),
                        SubModuleReplacementDescription("kv_b_proj", Linear1D_Col, kwargs=dict(**common)),
                        SubModuleReplacementDescription("o_proj", Linear1D_Row, kwargs=dict(**common)),
                    ])
            mlp_subs = [SubModuleReplacementDescription("down_proj", Linear1D_Row, kwargs=dict(**common))]
            if cfg.glu:
                mlp_subs.insert(0, SubModuleReplacementDescription(
                    "gate_up_proj", FusedLinear1D_Col, kwargs=dict(split_sizes=None, **common)))
            else:
                mlp_subs.insert(0, SubModuleReplacementDescription("up_proj", Linear1D_Col, kwargs=dict(**common)))
            policy[MLP] = ModulePolicyDescription(sub_module_replacement=mlp_subs, param_replacement=[_fix_glu_split])
            policy[TransformerModel] = ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription(
                    "embed_tokens", VocabParallelEmbedding1D,
                    kwargs=dict(make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                                fp8_communication=sc.fp8_communication))])
            policy[TransformerLMHeadModel] = ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription(
                    "lm_head", VocabParallelLMHead1D,
                    kwargs=dict(gather_output=not sc.parallel_output,
                                make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                                seq_parallel_mode=lin_sp, seq_parallel_dim=0,
                                fp8_communication=sc.fp8_communication))])
        elif zbv:
            policy[Attention] = ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription("qkv_proj", LinearWithGradAccum, kwargs=dict(use_zbv=True)),
                SubModuleReplacementDescription("o_proj", LinearWithGradAccum, kwargs=dict(use_zbv=True))])
            subs = [SubModuleReplacementDescription("down_proj", LinearWithGradAccum, kwargs=dict(use_zbv=True))]
            subs.append(SubModuleReplacementDescription("gate_up_proj" if cfg.glu else "up_proj", LinearWithGradAccum,
                                                        kwargs=dict(use_zbv=True)))
            policy[MLP] = ModulePolicyDescription(sub_module_replacement=subs)

        # norms: same module, but mark params whose grads are partial under split_gather / ring SP
        if sp_partial:
            norm_desc = [
                SubModuleReplacementDescription("input_layernorm", norm_cls, kwargs=dict(sp_partial_derived=True)),
                SubModuleReplacementDescription("post_attention_layernorm", norm_cls,
                                                kwargs=dict(sp_partial_derived=True), ignore_if_not_exist=True),
            ]
            self.append_or_create_submodule_replacement(norm_desc, policy, DecoderLayer)
            self.append_or_create_submodule_replacement(
                [SubModuleReplacementDescription("norm", norm_cls, kwargs=dict(sp_partial_derived=True),
                                                 ignore_if_not_exist=True)], policy, TransformerModel)
        # expert parallel blocks
        if cfg.moe is not None:
            policy.setdefault(SparseMoE, ModulePolicyDescription()).param_replacement = [
                lambda m, sc=sc: m.setup_parallel(sc)]
        return policy

    def postprocess(self) -> nn.Module:
        sc = self.shard_config
        for m in self.model.modules():
            if hasattr(m, "shard_config"):
                m.shard_config = sc
        # re-tie embeddings after both sides were sharded identically (vocab dim 0)
        cfg = self.model.cfg
        if cfg.tie_word_embeddings and hasattr(self.model, "lm_head") and self.model.lm_head is not None:
            emb = self.model.model.embed_tokens
            head = self.model.lm_head
            if emb is not None and getattr(emb, "weight", None) is not None and getattr(head, "weight", None) is not None \
                    and emb.weight.shape == head.weight.shape:
                head.weight = emb.weight
        return self.model

    # ------------------------------------------------------------------ pipeline
    def _backbone(self):
        return self.model.model if hasattr(self.model, "model") else self.model

    def get_held_layers(self) -> List[nn.Module]:
        sm = self.pipeline_stage_manager
        backbone = self._backbone()
        if sm is None:
            return [self.model]
        layers_per = sm.distribute_layers(len(backbone.layers))
        held: List[nn.Module] = []
        if sm.is_interleave:
            ranges = sm.get_stage_index(layers_per)
            ranges = ranges if isinstance(ranges, list) else [ranges]
            first = sm.is_first_stage(ignore_chunk=True)
            last_holder = (sm.stage == 0) if sm.use_zbv else sm.is_last_stage(ignore_chunk=True)
            for s, e in ranges:
                held.extend(backbone.layers[s:e])
        else:
            s, e = sm.get_stage_index(layers_per)
            held.extend(backbone.layers[s:e])
            first, last_holder = sm.is_first_stage(), sm.is_last_stage()
        if first:
            for n in ("embed_tokens", "embed_positions", "embed_token_types", "embed_layernorm"):
                if hasattr(backbone, n):
                    held.append(getattr(backbone, n))
        if last_holder:
            if hasattr(backbone, "norm"):
                held.append(backbone.norm)
            if hasattr(self.model, "lm_head") and self.model.lm_head is not None:
                held.append(self.model.lm_head)
            for n in ("score", "classifier", "pooler", "cls_head"):
                if hasattr(self.model, n):
                    held.append(getattr(self.model, n))
        return held

    def get_shared_params(self) -> List[Dict[int, Tensor]]:
        sm = self.pipeline_stage_manager
        cfg = self.model.cfg
        if sm is None or sm.num_stages == 1 or not cfg.tie_word_embeddings or not hasattr(self.model, "lm_head"):
            return []
        emb_w = self._backbone().embed_tokens.weight
        head_w = self.model.lm_head.weight
        if id(emb_w) == id(head_w):
            last = 0 if (sm.is_interleave and sm.use_zbv) else sm.num_stages - 1
            if last != 0:
                return [{0: emb_w, last: head_w}]
        return []


def _fix_glu_split(mlp: nn.Module) -> None:
    """no-op hook kept so MLP policies can attach per-module fixes (split sizes are resolved at replacement)."""
    return None


class TransformerForCausalLMPolicy(TransformerPolicy):
    pass
