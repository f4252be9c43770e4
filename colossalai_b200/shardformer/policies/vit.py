"""Sharding policies for the ViT family.  Parity: reference `colossalai/shardformer/policies/vit.py:24-290`
(`ViTModelPolicy`, `ViTForImageClassificationPolicy`, `ViTForMaskedImageModelingPolicy`): TP over heads / MLP width,
pipeline stages over the encoder layers (embeddings on the first stage, final norm / pooler / head on the last)."""
from __future__ import annotations

from typing import List

import torch.nn as nn

from .encdec import EncDecPolicy

__all__ = ["ViTModelPolicy", "ViTForImageClassificationPolicy", "ViTForMaskedImageModelingPolicy"]


class ViTModelPolicy(EncDecPolicy):
    head_fields = ["num_attention_heads"]

    def _vit(self) -> nn.Module:
        return self.model.vit if hasattr(self.model, "vit") else self.model

    def get_held_layers(self) -> List[nn.Module]:
        sm = self.pipeline_stage_manager
        if sm is None:
            return [self.model]
        vit = self._vit()
        s, e = sm.get_stage_index(sm.distribute_layers(len(vit.layers)))
        held: List[nn.Module] = list(vit.layers[s:e])
        if sm.is_first_stage():
            held.append(vit.embeddings)
        if sm.is_last_stage():
            held.append(vit.layernorm)
            if vit.pooler is not None:
                held.append(vit.pooler)
            for n in ("classifier", "decoder"):
                if hasattr(self.model, n):
                    held.append(getattr(self.model, n))
        return held


class ViTForImageClassificationPolicy(ViTModelPolicy):
    pass


class ViTForMaskedImageModelingPolicy(ViTModelPolicy):
    pass
