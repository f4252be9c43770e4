"""Policies that shard a user's HuggingFace SAM / BLIP-2 model in place (reference `policies/sam.py:1-210`,
`policies/blip2.py:20-420`, `modeling/{sam,blip2}.py`).

SAM.  Vision encoder: the fused `qkv` linear ([q | k | v], each head-major) is split per projection so that a rank keeps
the q, k and v columns of ITS heads (`FusedLinear1D_Col`), `proj` is row-parallel, the MLP block (`lin1` / `lin2`) column /
row; the module reshapes with `num_attention_heads`, which becomes the local count.  The decomposed relative-position
tables (`rel_pos_h`, `rel_pos_w`: [2 L - 1, head_dim], shared by all heads) stay replicated, but every rank only sees the
contribution of its own heads to their gradient, so a tensor hook sums that gradient over the TP group.  Mask decoder:
`SamAttention` (`q_proj` / `k_proj` / `v_proj` column, `out_proj` row, local head count) in the two-way transformer and its
final token-to-image attention.  Prompt encoder, neck, upscaling and the hyper-network MLPs are small and stay
replicated.

BLIP-2.  Vision tower: `Blip2Attention` computes its head width from the INPUT width and the head count, so its forward
is rebound to one that uses the module's `head_dim` (the only method replacement here); fused `qkv` split per projection,
`projection` row-parallel, `Blip2MLP` column / row.  Q-Former: BERT-style `query` / `key` / `value` column (self- and
cross-attention), `Blip2QFormerSelfOutput.dense` row, intermediate / output (text and query branches) column / row, with
the local `num_attention_heads` / `all_head_size`.  Language model: whatever in-place policy the auto lookup has for its
class (OPT, T5, Llama, ...) is merged in.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ..layer import Linear1D_Col, Linear1D_Row
from ..layer.qkv_fused_linear import FusedLinear1D_Col
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription

__all__ = ["HFSamPolicy", "HFBlip2Policy"]


class _HFVisionPolicy(Policy):
    def preprocess(self) -> nn.Module:
        return self.model

    def postprocess(self) -> nn.Module:
        return self.model

    def get_held_layers(self) -> List[nn.Module]:
        if self.pipeline_stage_manager is not None:
            raise NotImplementedError("pipeline parallelism of HuggingFace vision models: import the weights into the "
                                      "native zoo (`models.hf_io`) and use its policy")
        return []

    def get_shared_params(self):
        return []

    def _check_no_sp(self) -> None:
        assert not self.shard_config.enable_sequence_parallelism, \
            "sequence parallelism of HF vision modules is not supported; build the model from the native zoo"


def _sum_grad_over(group):
    """Parameter-replacement hook factory: replicated tables used by head-sharded attention get the sum of every rank's
    partial gradient (a tensor hook: runs on each backward's gradient, so gradient accumulation stays correct)."""

    def install(module: nn.Module) -> None:
        for name in ("rel_pos_h", "rel_pos_w"):
            p = getattr(module, name, None)
            if isinstance(p, nn.Parameter) and p.requires_grad and not getattr(p, "_cb200_tp_summed", False):
                def hook(g, _group=group):
                    g = g.contiguous().clone()
                    dist.all_reduce(g, group=_group)
                    return g

                p.register_hook(hook)
                p._cb200_tp_summed = True
    return install


class HFSamPolicy(_HFVisionPolicy):
    """`SamModel` (and `SamVisionModel`)."""

    def config_sanity_check(self) -> None:
        self._check_no_sp()
        if self.shard_config.enable_tensor_parallelism:
            cfg, tp = self.model.config, self.shard_config.tensor_parallel_size
            vis = getattr(cfg, "vision_config", cfg)
            assert vis.num_attention_heads % tp == 0, "vision num_attention_heads must be divisible by the TP size"
            dec = getattr(cfg, "mask_decoder_config", None)
            if dec is not None:
                assert dec.num_attention_heads % tp == 0, "mask-decoder num_attention_heads must be divisible by TP"

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        vis = getattr(cfg, "vision_config", cfg)
        fp8 = dict(fp8_communication=sc.fp8_communication)
        vision_attn = ModulePolicyDescription(
            attribute_replacement={"num_attention_heads": vis.num_attention_heads // tp},
            sub_module_replacement=[
                SubModuleReplacementDescription("qkv", FusedLinear1D_Col, kwargs=dict(fp8, num_splits=3)),
                SubModuleReplacementDescription("proj", Linear1D_Row, kwargs=dict(fp8))],
            param_replacement=[_sum_grad_over(sc.tensor_parallel_process_group)])
        policy["SamVisionAttention"] = vision_attn                  # (`attn_implementation="eager"`)
        policy["SamVisionSdpaAttention"] = vision_attn
        policy["SamMLPBlock"] = ModulePolicyDescription(sub_module_replacement=[   # vision layers AND two-way blocks
            SubModuleReplacementDescription("lin1", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("lin2", Linear1D_Row, kwargs=dict(fp8))])
        dec = getattr(cfg, "mask_decoder_config", None)
        if dec is not None:
            policy["SamAttention"] = ModulePolicyDescription(
                attribute_replacement={"num_attention_heads": dec.num_attention_heads // tp},
                sub_module_replacement=[
                    SubModuleReplacementDescription("q_proj", Linear1D_Col, kwargs=dict(fp8)),
                    SubModuleReplacementDescription("k_proj", Linear1D_Col, kwargs=dict(fp8)),
                    SubModuleReplacementDescription("v_proj", Linear1D_Col, kwargs=dict(fp8)),
                    SubModuleReplacementDescription("out_proj", Linear1D_Row, kwargs=dict(fp8))])
        return policy


def _blip2_attention_forward(self, hidden_states: torch.Tensor, **kwargs):
    """`Blip2Attention.forward` with the head width taken from the module (`head_dim`) instead of
    `input_width // num_heads` (wrong once the fused qkv holds only this rank's heads)."""
    bsz, tgt_len, _ = hidden_states.size()
    qkv = self.qkv(hidden_states).reshape(bsz, tgt_len, 3, -1, self.head_dim).permute(2, 0, 3, 1, 4)
    out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], scale=self.scale,
                                         dropout_p=self.attention_dropout if self.training else 0.0)
    out = out.transpose(1, 2).reshape(bsz, tgt_len, -1)
    return self.projection(out), None


class HFBlip2Policy(_HFVisionPolicy):
    """`Blip2Model`, `Blip2ForConditionalGeneration`."""

    def _lm_policy(self) -> Optional[Policy]:
        lm = getattr(self.model, "language_model", None)
        if lm is None:
            return None
        if getattr(self, "_lm", None) is None:
            from .auto_policy import get_autopolicy

            pol = get_autopolicy(lm)
            pol.set_model(lm)
            pol.set_shard_config(self.shard_config)
            self._lm = pol
        return self._lm

    def config_sanity_check(self) -> None:
        self._check_no_sp()
        if self.shard_config.enable_tensor_parallelism:
            cfg, tp = self.model.config, self.shard_config.tensor_parallel_size
            assert cfg.vision_config.num_attention_heads % tp == 0 and cfg.qformer_config.num_attention_heads % tp == 0, \
                "vision and Q-Former num_attention_heads must be divisible by the TP size"
        lm = self._lm_policy()
        if lm is not None:
            lm.config_sanity_check()

    def preprocess(self) -> nn.Module:
        lm = self._lm_policy()
        if lm is not None:
            lm.preprocess()
        return self.model

    def postprocess(self) -> nn.Module:
        lm = self._lm_policy()
        if lm is not None:
            lm.postprocess()
        return self.model

    def module_policy(self) -> Dict[str, ModulePolicyDescription]:
        sc = self.shard_config
        policy: Dict[str, ModulePolicyDescription] = {}
        lm = self._lm_policy()
        if lm is not None:
            policy.update(lm.module_policy())
        if not sc.enable_tensor_parallelism:
            return policy
        cfg, tp = self.model.config, sc.tensor_parallel_size
        fp8 = dict(fp8_communication=sc.fp8_communication)
        policy["Blip2Attention"] = ModulePolicyDescription(
            attribute_replacement={"num_heads": cfg.vision_config.num_attention_heads // tp},
            sub_module_replacement=[
                SubModuleReplacementDescription("qkv", FusedLinear1D_Col, kwargs=dict(fp8, num_splits=3)),
                SubModuleReplacementDescription("projection", Linear1D_Row, kwargs=dict(fp8))],
            method_replacement={"forward": _blip2_attention_forward})
        policy["Blip2MLP"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("fc1", Linear1D_Col, kwargs=dict(fp8)),
            SubModuleReplacementDescription("fc2", Linear1D_Row, kwargs=dict(fp8))])
        q = cfg.qformer_config
        policy["Blip2QFormerMultiHeadAttention"] = ModulePolicyDescription(
            attribute_replacement={"num_attention_heads": q.num_attention_heads // tp,
                                   "all_head_size": q.hidden_size // tp},
            sub_module_replacement=[
                SubModuleReplacementDescription("query", Linear1D_Col, kwargs=dict(fp8)),
                SubModuleReplacementDescription("key", Linear1D_Col, kwargs=dict(fp8)),
                SubModuleReplacementDescription("value", Linear1D_Col, kwargs=dict(fp8))])
        policy["Blip2QFormerSelfOutput"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("dense", Linear1D_Row, kwargs=dict(fp8))])
        policy["Blip2QFormerIntermediate"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("dense", Linear1D_Col, kwargs=dict(fp8))])
        policy["Blip2QFormerOutput"] = ModulePolicyDescription(sub_module_replacement=[
            SubModuleReplacementDescription("dense", Linear1D_Row, kwargs=dict(fp8))])
        return policy
