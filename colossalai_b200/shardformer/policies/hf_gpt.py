"""Policies that shard a user's HuggingFace GPT-2 / OPT model in place (tensor parallelism by sub-module and attribute
replacement; reference `policies/gpt2.py:26-190`, `policies/opt.py:27-160`).

These two families need more than the llama-like decoders of `hf_decoder.py`:
  * GPT-2 keeps q|k|v in ONE `Conv1D` ([in, 3 * hidden] weight): the column shard has to take each rank's slice of q, of
    k and of v (`GPT2FusedLinearConv1D_Col` with split sizes), and `GPT2Attention.split_size` - the width the module
    uses to cut the fused output apart again - becomes the local width;
  * OPT's attention views its projections with `self.num_heads`, so that attribute (and `embed_dim`) is replaced by the
    local value;
  * both tie the LM head to the token embedding: both sides are sharded along the vocabulary identically and re-tied;
    the learned position embeddings stay replicated.
LayerNorms are left to PyTorch (replicated parameters, tiny).  `split_gather` sequence parallelism works as in
`hf_decoder.py` (`sequence_parallel_hooks`: the blocks run on sequence shards, one gather in front of each attention /
MLP module, row linears reduce-scatter; OPT has no MLP module and flattens [batch, seq] in its layer, so `fc1` / `fc2`
gather and scatter the rows themselves along dim 0 - the MLP is row-wise, the row order does not matter).  Pipeline
stages (1F1B) come from `HFDecoderPipelineMixin`: these backbones ADD learned positions to `inputs_embeds` (GPT-2, OPT),
or apply embedding dropout (GPT-2, GPT-J) or an embedding LayerNorm (BLOOM) to it, so on the stages behind the first
one those modules are replaced by neutral elements; the tied head on the last stage is kept in step with the first
stage's embedding by the plugin's shared-parameter all-reduce."""
from __future__ import annotations

from typing import Dict, List

import torch.nn as nn

from ..layer import Linear1D_Col, Linear1D_Row, VocabParallelEmbedding1D, VocabParallelLMHead1D
from ..layer.qkv_fused_linear import GPT2FusedLinearConv1D_Col, GPT2FusedLinearConv1D_Row
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription
from .hf_decoder import HFDecoderPipelineMixin

__all__ = ["HFGPT2Policy", "HFOPTPolicy", "HFGPTJPolicy", "HFBloomPolicy", "HFFalconPolicy"]


class _HFTiedDecoderPolicy(HFDecoderPipelineMixin, Policy):
    # pipeline stages (`HFDecoderPipelineMixin`): GPT-2 naming by default
    _pp_backbone, _pp_layers, _pp_final, _pp_every = ("transformer",), "h", "ln_f", ()
    _pp_first, _pp_neutral = ("wte", "wpe"), {"wpe": "zero", "drop": "identity"}

    def config_sanity_check(self) -> None:
        cfg = self.model.config
        tp = self.shard_config.tensor_parallel_size
        if self.shard_config.enable_tensor_parallelism:
            heads = getattr(cfg, "num_attention_heads", None) or cfg.n_head
            assert heads % tp == 0, "the number of attention heads must be divisible by the TP size"
        if self.shard_config.enable_sequence_parallelism:
            assert self.shard_config.sequence_parallelism_mode == "split_gather" and \
                self.shard_config.enable_tensor_parallelism, (
                    "HF modules support sequence parallelism in `split_gather` mode (with tensor parallelism); the "
                    "all_to_all / ring_attn modes need the native zoo (models.hf_io)")

    # where the blocks live and what they are made of (`split_gather` sequence parallelism, see `hf_decoder.py`)
    _sp_layers, _sp_attn, _sp_mlp, _sp_norms = "h", "attn", "mlp", ("ln_1", "ln_2")

    def _sp_kwargs(self):
        """(column kwargs, row kwargs, backbone hook list or None)."""
        sc = self.shard_config
        col = dict(fp8_communication=sc.fp8_communication)
        row = dict(col)
        if not sc.enable_sequence_parallelism:
            return col, row, None
        col.update(seq_parallel_mode="pre_gathered")
        row.update(seq_parallel_mode="split_gather", seq_parallel_dim=1)
        from .hf_decoder import sequence_parallel_hooks

        return col, row, [sequence_parallel_hooks(sc.tensor_parallel_process_group, self._sp_layers, self._sp_attn,
                                                  self._sp_mlp, self._sp_norms)]

    def preprocess(self) -> nn.Module:
        self.tie_weight = self.tie_weight_check()
        return self.model

    def postprocess(self) -> nn.Module:
        sm = self.pipeline_stage_manager
        single_stage = sm is None or sm.num_stages == 1
        if getattr(self, "tie_weight", False) and self.shard_config.enable_tensor_parallelism and single_stage:
            emb, head = self.model.get_input_embeddings(), self.model.get_output_embeddings()
            if head is not None and emb is not None and head.weight.shape == emb.weight.shape:
                head.weight = emb.weight
        self._install_pipeline_stage()
        return self.model

    def _vocab_kwargs(self) -> dict:
        sc = self.shard_config
        return dict(make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by, fp8_communication=sc.fp8_communication)


class HFGPT2Policy(_HFTiedDecoderPolicy):
    """`GPT2Model`, `GPT2LMHeadModel`."""

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        hidden = cfg.hidden_size
        inner = cfg.n_inner if getattr(cfg, "n_inner", None) is not None else 4 * hidden
        col, row, hooks = self._sp_kwargs()
        policy["GPT2Attention"] = ModulePolicyDescription(attribute_replacement={
            "embed_dim": hidden // tp, "split_size": hidden // tp, "num_heads": cfg.num_attention_heads // tp})
        policy["GPT2Block"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("attn.c_attn", GPT2FusedLinearConv1D_Col,
                                            kwargs=dict(split_sizes=[hidden] * 3, **col)),
            SubModuleReplacementDescription("attn.c_proj", GPT2FusedLinearConv1D_Row, kwargs=dict(row)),
            SubModuleReplacementDescription("mlp.c_fc", GPT2FusedLinearConv1D_Col,
                                            kwargs=dict(split_sizes=[inner], **col)),
            SubModuleReplacementDescription("mlp.c_proj", GPT2FusedLinearConv1D_Row, kwargs=dict(row)),
        ])
        policy["GPT2Model"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("wte", VocabParallelEmbedding1D, kwargs=self._vocab_kwargs())],
            param_replacement=hooks)
        policy["GPT2LMHeadModel"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("lm_head", VocabParallelLMHead1D,
                                            kwargs=dict(gather_output=True, **self._vocab_kwargs()))])
        return policy


class HFOPTPolicy(_HFTiedDecoderPolicy):
    """`OPTModel`, `OPTForCausalLM` (models without `project_in / project_out`, i.e. word_embed_proj_dim == hidden)."""

    # the OPT layer flattens [batch, seq] before `fc1`: its MLP is row-wise, so fc1 / fc2 gather and scatter those rows
    # themselves (dim 0, any row order) instead of through a hook on an MLP module (there is none)
    _sp_layers, _sp_attn, _sp_mlp, _sp_norms = "layers", "self_attn", None, ("self_attn_layer_norm", "final_layer_norm")
    _pp_backbone, _pp_layers, _pp_final = ("model", "decoder"), "layers", "final_layer_norm"
    _pp_first, _pp_neutral = ("embed_tokens", "embed_positions"), {"embed_positions": "zero"}

    def config_sanity_check(self) -> None:
        super().config_sanity_check()
        cfg = self.model.config
        assert getattr(cfg, "word_embed_proj_dim", cfg.hidden_size) == cfg.hidden_size, \
            "OPT variants with project_in / project_out (word_embed_proj_dim != hidden_size) are not supported"

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        col, row, hooks = self._sp_kwargs()
        flat = dict(row, seq_parallel_dim=0) if hooks else dict(row)
        policy["OPTAttention"] = ModulePolicyDescription(attribute_replacement={
            "embed_dim": cfg.hidden_size // tp, "num_heads": cfg.num_attention_heads // tp})
        policy["OPTDecoderLayer"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("self_attn.q_proj", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("self_attn.k_proj", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("self_attn.v_proj", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("self_attn.out_proj", Linear1D_Row, kwargs=dict(row)),
            SubModuleReplacementDescription("fc1", Linear1D_Col, kwargs=dict(flat)),
            SubModuleReplacementDescription("fc2", Linear1D_Row, kwargs=dict(flat)),
        ])
        policy["OPTDecoder"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("embed_tokens", VocabParallelEmbedding1D, kwargs=self._vocab_kwargs())],
            param_replacement=hooks)
        policy["OPTForCausalLM"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("lm_head", VocabParallelLMHead1D,
                                            kwargs=dict(gather_output=True, **self._vocab_kwargs()))])
        return policy


class HFGPTJPolicy(_HFTiedDecoderPolicy):
    """`GPTJModel`, `GPTJForCausalLM` (parallel attention + MLP block, partial rotary inside the attention module: the
    rotary slice is per head, so splitting heads over ranks leaves it untouched).  The LM head is not tied and has a
    bias; it becomes a gathered vocab-parallel head."""

    _sp_layers, _sp_attn, _sp_mlp, _sp_norms = "h", "attn", "mlp", ("ln_1",)
    _pp_first, _pp_neutral = ("wte",), {"drop": "identity"}

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        col, row, hooks = self._sp_kwargs()
        policy["GPTJAttention"] = ModulePolicyDescription(attribute_replacement={
            "embed_dim": cfg.hidden_size // tp, "num_attention_heads": cfg.num_attention_heads // tp})
        policy["GPTJBlock"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("attn.q_proj", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("attn.k_proj", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("attn.v_proj", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("attn.out_proj", Linear1D_Row, kwargs=dict(row)),
            SubModuleReplacementDescription("mlp.fc_in", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("mlp.fc_out", Linear1D_Row, kwargs=dict(row)),
        ])
        policy["GPTJModel"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("wte", VocabParallelEmbedding1D, kwargs=self._vocab_kwargs())],
            param_replacement=hooks)
        policy["GPTJForCausalLM"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("lm_head", VocabParallelLMHead1D,
                                            kwargs=dict(gather_output=True, **self._vocab_kwargs()))])
        return policy


def _bloom_attention_forward(self, hidden_states, residual, alibi, attention_mask, *args, **kwargs):
    """`BloomModel` builds the ALiBi tensor for ALL heads ([batch * heads, 1, kv]); a tensor-parallel attention block
    owns a contiguous range of heads, so it takes its slice and runs the module's own forward."""
    tp, rank = self._cb200_tp_size, self._cb200_tp_rank
    batch = hidden_states.shape[0]
    local = self.num_heads
    alibi = alibi.view(batch, local * tp, *alibi.shape[1:])[:, rank * local:(rank + 1) * local]
    alibi = alibi.reshape(batch * local, *alibi.shape[2:])
    return type(self).forward(self, hidden_states, residual, alibi, attention_mask, *args, **kwargs)


class HFBloomPolicy(_HFTiedDecoderPolicy):
    """`BloomModel`, `BloomForCausalLM` (reference `policies/bloom.py:25-240`).  The fused `query_key_value` weight is
    laid out [heads, 3, head_dim], i.e. head-major: a contiguous row split is a split by heads, so the plain column
    linear applies; `num_heads` / `hidden_size` become local values and the attention forward is wrapped to slice the
    ALiBi bias of its heads.  Assumes `pretraining_tp == 1` (the HF default for fine-tuning)."""

    _sp_layers, _sp_attn, _sp_mlp, _sp_norms = "h", "self_attention", "mlp", ("input_layernorm", "post_attention_layernorm")
    _pp_first, _pp_neutral = ("word_embeddings", "word_embeddings_layernorm"), {"word_embeddings_layernorm": "identity"}

    def config_sanity_check(self) -> None:
        super().config_sanity_check()
        cfg = self.model.config
        assert getattr(cfg, "pretraining_tp", 1) == 1 or not getattr(cfg, "slow_but_exact", False), \
            "slow_but_exact with pretraining_tp > 1 slices the full hidden size inside the module"

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        from ...parallel import comm

        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        fp8 = dict(fp8_communication=sc.fp8_communication)
        rank = comm.group_rank(sc.tensor_parallel_process_group)
        policy["BloomAttention"] = ModulePolicyDescription(
            attribute_replacement={"num_heads": cfg.n_head // tp, "hidden_size": cfg.hidden_size // tp,
                                   "_cb200_tp_size": tp, "_cb200_tp_rank": rank},
            method_replacement={"forward": _bloom_attention_forward})
        col, row, hooks = self._sp_kwargs()
        policy["BloomBlock"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("self_attention.query_key_value", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("self_attention.dense", Linear1D_Row, kwargs=dict(row)),
            SubModuleReplacementDescription("mlp.dense_h_to_4h", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("mlp.dense_4h_to_h", Linear1D_Row, kwargs=dict(row)),
        ])
        policy["BloomModel"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("word_embeddings", VocabParallelEmbedding1D, kwargs=self._vocab_kwargs())],
            param_replacement=hooks)
        policy["BloomForCausalLM"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("lm_head", VocabParallelLMHead1D,
                                            kwargs=dict(gather_output=True, **self._vocab_kwargs()))])
        return policy


class HFFalconPolicy(_HFTiedDecoderPolicy):
    """`FalconModel`, `FalconForCausalLM` (reference `policies/falcon.py:24-200`) for the two fused-QKV layouts that
    are group-major and therefore split cleanly by rows: the new decoder architecture ([kv groups, q heads per group
    + 2, head_dim]: a row split is a split by kv groups) and the original multi-head layout ([heads, 3, head_dim]).
    The multi-query layout of Falcon-7B keeps ONE key / value head behind all query heads: splitting the query heads
    would leave K / V replicated inside a column-parallel weight, which this in-place policy does not do (use the native
    zoo: `models.hf_io` imports Falcon checkpoints with the KV head replicated per rank).  ALiBi variants (falcon-rw) are
    likewise left to the native zoo."""

    _sp_layers, _sp_attn, _sp_mlp = "h", "self_attention", "mlp"
    _sp_norms = ("input_layernorm", "post_attention_layernorm", "ln_attn", "ln_mlp")
    _pp_first, _pp_neutral, _pp_every = ("word_embeddings",), {}, ("rotary_emb",)

    def config_sanity_check(self) -> None:
        super().config_sanity_check()
        cfg, tp = self.model.config, self.shard_config.tensor_parallel_size
        if not self.shard_config.enable_tensor_parallelism:
            return
        assert not getattr(cfg, "alibi", False), "Falcon with ALiBi: use the native zoo (models.hf_io)"
        if cfg.new_decoder_architecture:
            assert cfg.num_kv_heads % tp == 0, "num_kv_heads must be divisible by the TP size"
        else:
            assert not cfg.multi_query, "multi-query Falcon (one shared KV head): use the native zoo (models.hf_io)"

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        fp8 = dict(fp8_communication=sc.fp8_communication)
        attrs = {"num_heads": cfg.num_attention_heads // tp, "hidden_size": cfg.hidden_size // tp,
                 "split_size": cfg.hidden_size // tp}
        if cfg.new_decoder_architecture:
            attrs["num_kv_heads"] = cfg.num_kv_heads // tp
        elif not cfg.multi_query:
            attrs["num_kv_heads"] = cfg.num_attention_heads // tp
        policy["FalconAttention"] = ModulePolicyDescription(attribute_replacement=attrs)
        col, row, hooks = self._sp_kwargs()
        policy["FalconDecoderLayer"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("self_attention.query_key_value", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("self_attention.dense", Linear1D_Row, kwargs=dict(row)),
            SubModuleReplacementDescription("mlp.dense_h_to_4h", Linear1D_Col, kwargs=dict(col)),
            SubModuleReplacementDescription("mlp.dense_4h_to_h", Linear1D_Row, kwargs=dict(row)),
        ])
        policy["FalconModel"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("word_embeddings", VocabParallelEmbedding1D, kwargs=self._vocab_kwargs())],
            param_replacement=hooks)
        policy["FalconForCausalLM"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("lm_head", VocabParallelLMHead1D,
                                            kwargs=dict(gather_output=True, **self._vocab_kwargs()))])
        return policy
