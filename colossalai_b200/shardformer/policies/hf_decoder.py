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
  * pipeline parallelism (1F1B, one chunk per stage) needs NO rewrite of the decoder's forward: the HF backbone accepts
    `inputs_embeds`, loops over `self.layers` and ends with `self.norm`, so a stage keeps its slice of `layers`
    (`embed_tokens` on the first stage, `norm` + `lm_head` on the last, the rotary table everywhere), `norm` becomes
    an identity on the other stages, and a small stage forward on the top-level module feeds `input_ids` (first stage)
    or the previous stage's `hidden_states` (as `inputs_embeds`) into the backbone and returns hidden states or
    logits + loss.  Tied embeddings across the first and last stage are reported through `get_shared_params`.
  * sequence parallelism, `split_gather` mode (Megatron SP inside the TP group): HF attention reshapes q/k/v with the
    sequence length of its INPUT, so the sequence is gathered once in front of each attention / MLP block (forward
    pre-hook: all-gather forward, reduce-scatter backward - one gather feeds q, k and v instead of one per linear),
    the column linears then run without communication (`pre_gathered`: their dX stays a partial sum for that
    reduce-scatter) and the row linears reduce-scatter their output along the sequence (`seq_parallel_dim=1`: HF
    activations are [batch, seq, hidden]).  The HF backbone computes rotary tables and the causal mask for the FULL
    sequence before its layer loop, so the only other plumbing is a split of the hidden states in front of the first
    decoder layer and a gather behind the last one.  The norms inside the layers see sequence shards, so their
    weights are marked SP-partial (the plugin all-reduces those gradients over the TP group).  The other SP modes
    (Ulysses, ring attention) need the attention rewritten: native zoo.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from ..layer import Linear1D_Col, Linear1D_Row, VocabParallelEmbedding1D, VocabParallelLMHead1D
from ..layer.qkv_fused_linear import FusedLinear1D_Col
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription

__all__ = ["HFDecoderPolicy", "HFDecoderPipelineMixin", "HF_FAMILIES", "sequence_parallel_hooks"]

# transformers module path -> class-name prefix
HF_FAMILIES = {"llama": "Llama", "mistral": "Mistral", "qwen2": "Qwen2", "qwen3": "Qwen3", "cohere": "Cohere",
               "glm": "Glm"}
# families whose MLP keeps gate and up in ONE linear (`gate_up_proj`, output = [gate | up]): each half is split over the
# ranks separately so that the module's own `chunk(2, -1)` still yields matching gate / up slices
FUSED_GATE_UP = ("Glm",)


def _pp_stage_forward(self, input_ids=None, hidden_states=None, labels=None, attention_mask=None, position_ids=None,
                      **kwargs):
    """Stage-aware forward bound to the user's `<Family>ForCausalLM` / `<Family>Model` under pipeline parallelism."""
    sm = self._cb200_stage_manager
    backbone = self._cb200_backbone
    inner = type(backbone).forward                       # the HF backbone's own forward
    kw = dict(attention_mask=attention_mask, position_ids=position_ids, use_cache=False)
    if sm.is_first_stage():
        kw["input_ids"] = input_ids
    else:
        kw["inputs_embeds"] = hidden_states
    accepted = self._cb200_backbone_params
    if accepted is not None:                             # (BLOOM / Falcon backbones take no `position_ids`)
        kw = {k: v for k, v in kw.items() if k in accepted}
    h = inner(backbone, **kw).last_hidden_state
    if not sm.is_last_stage():
        return {"hidden_states": h}
    if not hasattr(self, "lm_head"):
        return {"last_hidden_state": h}
    logits = self.lm_head(h)
    scale = getattr(self, "logit_scale", None)           # Cohere
    if scale is not None:
        logits = logits * scale
    loss = None
    if labels is not None:
        loss = self.loss_function(logits=logits, labels=labels, vocab_size=self.config.vocab_size)
    return {"loss": loss, "logits": logits}


class _ZeroPositions(nn.Module):
    """Stands in for a learned position embedding on the stages behind the first one: the HF backbone ADDS the position
    embedding to `inputs_embeds`, which on those stages are hidden states that already carry it."""

    def forward(self, *args, **kwargs) -> torch.Tensor:
        return torch.zeros(())


def _fused_rmsnorm_forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
    from ... import ops

    eps = getattr(self, "variance_epsilon", getattr(self, "eps", 1e-6))
    weight = self.weight
    group = getattr(self, "_cb200_tp_group", None)
    if group is not None:
        # per-head q / k norm under tensor parallelism: the [head_dim] weight is replicated while every rank only sees
        # its own heads, so its gradient is a partial sum - identity forward, all-reduce backward over the TP group
        from ..layer._operation import reduce_backward

        weight = reduce_backward(weight, group)
    return ops.rms_norm(hidden_states, weight, eps)


def mark_head_norms(tp_group):
    """Parameter-replacement hook for `<Family>Attention`: q_norm / k_norm (Qwen3-style per-head RMSNorm) get the
    TP-aware forward above."""
    from types import MethodType

    def hook(attn: nn.Module) -> None:
        for name in ("q_norm", "k_norm"):
            norm = getattr(attn, name, None)
            if norm is not None and getattr(norm, "weight", None) is not None and norm.weight.dim() == 1:
                norm._cb200_tp_group = tp_group
                norm.forward = MethodType(_fused_rmsnorm_forward, norm)
    return hook


class HFDecoderPipelineMixin:
    """1F1B pipeline stages for HF decoders whose backbone accepts `inputs_embeds`: a stage keeps its slice of the
    block list and the HF loop walks just that.  The class attributes name the family's modules: the path from the LM
    wrapper to the backbone, the block list, the modules only the first stage needs, the final norm, modules every
    stage needs (rotary tables), and what the backbone applies to `inputs_embeds` BEFORE the first block - on the
    stages behind the first one those are replaced by neutral elements (learned positions -> zeros, embedding dropout /
    embedding LayerNorm -> identity), because their input is a hidden state, not an embedding."""

    _pp_backbone = ("model",)
    _pp_layers = "layers"
    _pp_first = ("embed_tokens",)
    _pp_final = "norm"
    _pp_every = ("rotary_emb",)
    _pp_neutral = {}                     # attribute -> "zero" | "identity"

    def _backbone(self) -> nn.Module:
        m = self.model
        if hasattr(m, "lm_head"):
            for attr in self._pp_backbone:
                m = getattr(m, attr)
        return m

    def _stage_layers(self):
        sm = self.pipeline_stage_manager
        layers = getattr(self._backbone(), self._pp_layers)
        assert not sm.is_interleave, "HF modules support the 1F1B schedule (one model chunk per stage)"
        start, end = sm.get_stage_index(sm.distribute_layers(len(layers)))
        return list(layers[start:end])

    def _install_pipeline_stage(self) -> None:
        sm = self.pipeline_stage_manager
        if sm is None or sm.num_stages == 1:
            return
        import inspect
        from types import MethodType

        backbone = self._backbone()
        setattr(backbone, self._pp_layers, nn.ModuleList(self._held_decoder_layers))   # the HF loop walks this stage only
        if not sm.is_last_stage():
            setattr(backbone, self._pp_final, nn.Identity())
        if not sm.is_first_stage():
            for attr, kind in self._pp_neutral.items():
                if hasattr(backbone, attr):
                    setattr(backbone, attr, _ZeroPositions() if kind == "zero" else nn.Identity())
        params = inspect.signature(type(backbone).forward).parameters
        var_kw = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())
        self.model._cb200_backbone_params = None if var_kw else set(params)
        self.model._cb200_backbone = backbone
        self.model._cb200_stage_manager = sm
        self.model.forward = MethodType(_pp_stage_forward, self.model)

    def get_held_layers(self) -> List[nn.Module]:
        sm = self.pipeline_stage_manager
        if sm is None:
            return []
        backbone = self._backbone()
        self._held_decoder_layers = self._stage_layers()
        held: List[nn.Module] = list(self._held_decoder_layers)
        held += [getattr(backbone, a) for a in self._pp_every if hasattr(backbone, a)]
        if sm.is_first_stage():
            held += [getattr(backbone, a) for a in self._pp_first if hasattr(backbone, a)]
        if sm.is_last_stage():
            held.append(getattr(backbone, self._pp_final))
            if hasattr(self.model, "lm_head"):
                held.append(self.model.lm_head)
        return held

    def get_shared_params(self):
        sm = self.pipeline_stage_manager
        if sm is None or sm.num_stages == 1 or not getattr(self, "tie_weight", False) or not hasattr(self.model, "lm_head"):
            return []
        emb_w, head_w = self.model.get_input_embeddings().weight, self.model.lm_head.weight
        return [{0: emb_w, sm.num_stages - 1: head_w}]


def sequence_parallel_hooks(group, layers_attr: str = "layers", attn_attr: Optional[str] = "self_attn",
                            mlp_attr: Optional[str] = "mlp", norm_attrs=("input_layernorm", "post_attention_layernorm",
                                                                         "pre_feedforward_layernorm",
                                                                         "post_feedforward_layernorm")):
    """Parameter-replacement hook for the backbone module of an HF decoder under `split_gather` sequence parallelism:
    split the sequence in front of the first block, gather it behind the last one (autograd-aware), gather ONCE in front
    of every attention / MLP sub-module (all-gather forward, reduce-scatter backward; only their first argument - BLOOM
    passes the sequence-sharded residual as the second), mark the in-block norm parameters SP-partial."""
    from ..layer._operation import (gather_forward_reducescatter_backward, gather_forward_split_backward,
                                    split_forward_gather_backward)
    from ..layer.utils import SeqParallelUtils

    def first_arg(fn):
        def hook(module, args, kwargs):
            if args:
                return (fn(args[0]),) + tuple(args[1:]), kwargs
            kwargs = dict(kwargs)
            kwargs["hidden_states"] = fn(kwargs["hidden_states"])
            return args, kwargs
        return hook

    def gather_out(module, args, output):
        if isinstance(output, tuple):
            return (gather_forward_split_backward(output[0], 1, group),) + tuple(output[1:])
        return gather_forward_split_backward(output, 1, group)

    def install(backbone: nn.Module) -> None:
        layers = list(getattr(backbone, layers_attr))
        if not layers:
            return
        layers[0].register_forward_pre_hook(first_arg(lambda t: split_forward_gather_backward(t, 1, group)),
                                            with_kwargs=True)
        layers[-1].register_forward_hook(gather_out)
        gather_in = first_arg(lambda t: gather_forward_reducescatter_backward(t, group, 1))
        for layer in layers:
            for attr in (attn_attr, mlp_attr):
                if attr is not None:
                    getattr(layer, attr).register_forward_pre_hook(gather_in, with_kwargs=True)
            for name in norm_attrs:
                norm = getattr(layer, name, None)
                if norm is not None:
                    for prm in norm.parameters(recurse=False):
                        SeqParallelUtils.marked_as_sp_partial_derived_param(prm)
    return install


class HFDecoderPolicy(HFDecoderPipelineMixin, Policy):
    """Works for `<Family>Model`, `<Family>ForCausalLM` of the families in HF_FAMILIES (matched by class NAME, so the
    policy never imports transformers itself)."""

    def config_sanity_check(self) -> None:
        cfg = self.model.config
        tp = self.shard_config.tensor_parallel_size
        if self.shard_config.enable_tensor_parallelism:
            assert cfg.num_attention_heads % tp == 0, "num_attention_heads must be divisible by the TP size"
            kv = getattr(cfg, "num_key_value_heads", cfg.num_attention_heads)
            assert kv % tp == 0, "num_key_value_heads must be divisible by the TP size"
        if self.shard_config.enable_sequence_parallelism:
            assert self.shard_config.sequence_parallelism_mode == "split_gather" and \
                self.shard_config.enable_tensor_parallelism, (
                    "HF modules support sequence parallelism in `split_gather` mode (with tensor parallelism); the "
                    "all_to_all / ring_attn modes need the native zoo (models.hf_io)")

    def preprocess(self) -> nn.Module:
        self.tie_weight = self.tie_weight_check()
        return self.model

    def _sp_hooks(self):
        return sequence_parallel_hooks(self.shard_config.tensor_parallel_process_group)

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
            row = dict(col)
            if sc.enable_sequence_parallelism:
                col.update(seq_parallel_mode="pre_gathered")
                row.update(seq_parallel_mode="split_gather", seq_parallel_dim=1)
            if fam in FUSED_GATE_UP:
                mlp_in = [SubModuleReplacementDescription("mlp.gate_up_proj", FusedLinear1D_Col,
                                                          kwargs=dict(col, num_splits=2))]
            else:
                mlp_in = [SubModuleReplacementDescription("mlp.gate_proj", Linear1D_Col, kwargs=dict(col)),
                          SubModuleReplacementDescription("mlp.up_proj", Linear1D_Col, kwargs=dict(col))]
            policy[f"{fam}DecoderLayer"] = ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription("self_attn.q_proj", Linear1D_Col, kwargs=dict(col)),
                SubModuleReplacementDescription("self_attn.k_proj", Linear1D_Col, kwargs=dict(col)),
                SubModuleReplacementDescription("self_attn.v_proj", Linear1D_Col, kwargs=dict(col)),
                SubModuleReplacementDescription("self_attn.o_proj", Linear1D_Row, kwargs=dict(row)),
                *mlp_in,
                SubModuleReplacementDescription("mlp.down_proj", Linear1D_Row, kwargs=dict(row)),
            ])
            policy[f"{fam}Attention"] = ModulePolicyDescription(
                param_replacement=[mark_head_norms(sc.tensor_parallel_process_group)])
            policy[f"{fam}Model"] = ModulePolicyDescription(
                sub_module_replacement=[SubModuleReplacementDescription(
                    "embed_tokens", VocabParallelEmbedding1D,
                    kwargs=dict(make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                                fp8_communication=sc.fp8_communication))],
                param_replacement=[self._sp_hooks()] if sc.enable_sequence_parallelism else None)
            policy[f"{fam}ForCausalLM"] = ModulePolicyDescription(sub_module_replacement=[
                SubModuleReplacementDescription(
                    "lm_head", VocabParallelLMHead1D,
                    kwargs=dict(gather_output=True, make_vocab_size_divisible_by=sc.make_vocab_size_divisible_by,
                                fp8_communication=sc.fp8_communication))])
        if sc.enable_fused_normalization:
            policy[f"{fam}RMSNorm"] = ModulePolicyDescription(method_replacement={"forward": _fused_rmsnorm_forward})
        return policy

    def postprocess(self) -> nn.Module:
        sm = self.pipeline_stage_manager
        single_stage = sm is None or sm.num_stages == 1 or (sm.is_first_stage() and sm.is_last_stage())
        if getattr(self, "tie_weight", False) and self.shard_config.enable_tensor_parallelism and single_stage:
            # both sides are sharded along the vocab dimension identically: re-tie the local shards
            emb = self.model.get_input_embeddings()
            head = self.model.get_output_embeddings()
            if head is not None and emb is not None and head.weight.shape == emb.weight.shape:
                head.weight = emb.weight
        self._install_pipeline_stage()
        return self.model
