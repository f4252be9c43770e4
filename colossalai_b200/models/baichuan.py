"""Baichuan / Baichuan-2 (RMSNorm, SwiGLU, fused `W_pack` QKV; RoPE for 7B, ALiBi for 13B; Baichuan-2's `NormHead`).

Parity: reference `colossalai/inference/modeling/models/nopadding_baichuan.py:1-420` +
`inference/modeling/policy/nopadding_baichuan.py` (the reference carries Baichuan on its inference path).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .config import ModelConfig, get_config
from .heads import TransformerBackboneModel, TransformerForSequenceClassification
from .transformer import TransformerLMHeadModel

DEFAULT_PRESET = "baichuan-7b"


def default_config(**overrides) -> ModelConfig:
    return get_config(DEFAULT_PRESET, **overrides)


class BaichuanModel(TransformerBackboneModel):
    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


class BaichuanForCausalLM(TransformerLMHeadModel):
    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)

    @torch.no_grad()
    def fold_norm_head(self) -> None:
        """Bake NormHead into the LM-head weight (inference: the rows never change, so normalise once)."""
        if self.cfg.norm_head:
            self.lm_head.weight.copy_(F.normalize(self.lm_head.weight.float(), dim=-1).to(self.lm_head.weight.dtype))
            self.cfg = self.config = self.cfg.replace(norm_head=False)
            self.model.cfg = self.cfg


class BaichuanForSequenceClassification(TransformerForSequenceClassification):
    def __init__(self, config: ModelConfig = None, **kw) -> None:
        super().__init__(config if config is not None else default_config(), **kw)


__all__ = ["default_config", "BaichuanModel", "BaichuanForCausalLM", "BaichuanForSequenceClassification"]
