"""Vision Transformer family: `ViTModel`, `ViTForImageClassification`, `ViTForMaskedImageModeling`.

Patch embedding is one strided conv (a GEMM over unfolded patches), the encoder is a stack of pre-norm
`EncoderBlock`s.  Pipeline-stage aware like the language models: the first stage owns the embeddings, the last the
final norm / pooler / head, and stages exchange `hidden_states`.

Parity: reference `colossalai/shardformer/policies/vit.py:24-290` + `modeling/vit.py:18-390` (the HF ViT classes).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encdec import BlockConfig, EncoderBlock, init_module_weights, make_norm

__all__ = ["ViTConfig", "ViTEmbeddings", "ViTModel", "ViTForImageClassification", "ViTForMaskedImageModeling",
           "VIT_ZOO"]


@dataclass
class ViTConfig:
    model_type: str = "vit"
    image_size: int = 224
    patch_size: int = 16
    num_channels: int = 3
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "gelu"
    layer_norm_eps: float = 1e-12
    qkv_bias: bool = True
    num_labels: int = 1000
    hidden_dropout: float = 0.0
    encoder_stride: int = 16
    initializer_range: float = 0.02
    add_pooling_layer: bool = True

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    def block(self) -> BlockConfig:
        return BlockConfig(hidden_size=self.hidden_size, num_heads=self.num_attention_heads,
                           ffn_dim=self.intermediate_size, act=self.hidden_act, qkv_bias=self.qkv_bias,
                           norm_eps=self.layer_norm_eps, pre_norm=True, dropout=self.hidden_dropout)

    def replace(self, **kw) -> "ViTConfig":
        return replace(self, **kw)


VIT_ZOO: Dict[str, ViTConfig] = {
    "vit-base": ViTConfig(),
    "vit-large": ViTConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096),
    "vit-tiny": ViTConfig(image_size=32, patch_size=8, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                          intermediate_size=128, num_labels=10, encoder_stride=8),
}


class ViTEmbeddings(nn.Module):
    def __init__(self, cfg: ViTConfig, use_mask_token: bool = False) -> None:
        super().__init__()
        self.cfg = cfg
        self.patch_embeddings = nn.Conv2d(cfg.num_channels, cfg.hidden_size, cfg.patch_size, stride=cfg.patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, cfg.hidden_size))
        self.position_embeddings = nn.Parameter(torch.zeros(1, cfg.num_patches + 1, cfg.hidden_size))
        self.mask_token = nn.Parameter(torch.zeros(1, 1, cfg.hidden_size)) if use_mask_token else None

    def forward(self, pixel_values: torch.Tensor, bool_masked_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self.patch_embeddings(pixel_values.to(self.patch_embeddings.weight.dtype)).flatten(2).transpose(1, 2)
        if bool_masked_pos is not None and self.mask_token is not None:
            m = bool_masked_pos.unsqueeze(-1).to(x.dtype)
            x = x * (1.0 - m) + self.mask_token.expand(x.shape[0], x.shape[1], -1) * m
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        pos = self.position_embeddings
        if pos.shape[1] != x.shape[1]:   # different resolution: bicubic interpolation of the patch grid
            n = int((x.shape[1] - 1) ** 0.5)
            g = int((pos.shape[1] - 1) ** 0.5)
            grid = pos[:, 1:].reshape(1, g, g, -1).permute(0, 3, 1, 2)
            grid = F.interpolate(grid, size=(n, n), mode="bicubic", align_corners=False)
            pos = torch.cat([pos[:, :1], grid.permute(0, 2, 3, 1).reshape(1, n * n, -1)], dim=1)
        x = x + pos
        if self.cfg.hidden_dropout > 0 and self.training:
            x = F.dropout(x, self.cfg.hidden_dropout)
        return x


class ViTModel(nn.Module):
    def __init__(self, config: Optional[ViTConfig] = None, add_pooling_layer: Optional[bool] = None,
                 use_mask_token: bool = False, _init: bool = True, **kw) -> None:
        super().__init__()
        cfg = config or ViTConfig(**kw)
        self.cfg = self.config = cfg
        self.embeddings = ViTEmbeddings(cfg, use_mask_token)
        bc = cfg.block()
        self.layers = nn.ModuleList([EncoderBlock(bc) for _ in range(cfg.num_hidden_layers)])
        self.layernorm = make_norm(bc)
        pool = cfg.add_pooling_layer if add_pooling_layer is None else add_pooling_layer
        self.pooler = nn.Linear(cfg.hidden_size, cfg.hidden_size) if pool else None
        self.shard_config = None
        self.gradient_checkpointing = False
        if _init:
            init_module_weights(self, cfg.initializer_range)

    def gradient_checkpointing_enable(self, *a, **k) -> None:
        self.gradient_checkpointing = True

    def _stage(self):
        sc = self.shard_config
        return sc.pipeline_stage_manager if sc is not None else None

    def layer_range(self):
        sm = self._stage()
        if sm is None:
            return 0, len(self.layers)
        return sm.get_stage_index(sm.distribute_layers(len(self.layers)))

    def forward(self, pixel_values: Optional[torch.Tensor] = None, hidden_states: Optional[torch.Tensor] = None,
                bool_masked_pos: Optional[torch.Tensor] = None, **unused) -> Dict[str, torch.Tensor]:
        sm = self._stage()
        first = sm is None or sm.is_first_stage()
        last = sm is None or sm.is_last_stage()
        x = self.embeddings(pixel_values, bool_masked_pos) if first else hidden_states
        s, e = self.layer_range()
        for i in range(s, e):
            if self.gradient_checkpointing and self.training:
                x = torch.utils.checkpoint.checkpoint(self.layers[i], x, use_reentrant=False)
            else:
                x = self.layers[i](x)
        if not last:
            return {"hidden_states": x}
        x = self.layernorm(x)
        out = {"last_hidden_state": x}
        if self.pooler is not None:
            out["pooler_output"] = torch.tanh(self.pooler(x[:, 0]))
        return out


class ViTForImageClassification(nn.Module):
    def __init__(self, config: Optional[ViTConfig] = None, **kw) -> None:
        super().__init__()
        cfg = config or ViTConfig(**kw)
        self.cfg = self.config = cfg
        self.vit = ViTModel(cfg, add_pooling_layer=False, _init=False)
        self.classifier = nn.Linear(cfg.hidden_size, cfg.num_labels)
        self.shard_config = None
        init_module_weights(self, cfg.initializer_range)

    def gradient_checkpointing_enable(self, *a, **k) -> None:
        self.vit.gradient_checkpointing = True

    def forward(self, pixel_values: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                hidden_states: Optional[torch.Tensor] = None, **unused) -> Dict[str, torch.Tensor]:
        out = self.vit(pixel_values=pixel_values, hidden_states=hidden_states)
        if "last_hidden_state" not in out:
            return out
        logits = self.classifier(out["last_hidden_state"][:, 0])
        res = {"logits": logits}
        if labels is not None:
            if labels.dtype in (torch.long, torch.int):
                res["loss"] = F.cross_entropy(logits.float(), labels)
            else:   # multi-label / regression targets
                res["loss"] = F.binary_cross_entropy_with_logits(logits.float(), labels.float())
        return res


class ViTForMaskedImageModeling(nn.Module):
    """SimMIM-style head: 1x1 conv to `stride^2 * C` channels + pixel shuffle, L1 loss on the masked patches."""

    def __init__(self, config: Optional[ViTConfig] = None, **kw) -> None:
        super().__init__()
        cfg = config or ViTConfig(**kw)
        self.cfg = self.config = cfg
        self.vit = ViTModel(cfg, add_pooling_layer=False, use_mask_token=True, _init=False)
        self.decoder = nn.Sequential(
            nn.Conv2d(cfg.hidden_size, cfg.encoder_stride ** 2 * cfg.num_channels, kernel_size=1),
            nn.PixelShuffle(cfg.encoder_stride))
        self.shard_config = None
        init_module_weights(self, cfg.initializer_range)

    def forward(self, pixel_values: Optional[torch.Tensor] = None, bool_masked_pos: Optional[torch.Tensor] = None,
                hidden_states: Optional[torch.Tensor] = None, **unused) -> Dict[str, torch.Tensor]:
        out = self.vit(pixel_values=pixel_values, hidden_states=hidden_states, bool_masked_pos=bool_masked_pos)
        if "last_hidden_state" not in out:
            return out
        seq = out["last_hidden_state"][:, 1:]
        B, N, C = seq.shape
        g = int(N ** 0.5)
        recon = self.decoder(seq.permute(0, 2, 1).reshape(B, C, g, g))
        res = {"reconstruction": recon, "logits": recon}
        if bool_masked_pos is not None and pixel_values is not None:
            cfg = self.cfg
            size = cfg.image_size // cfg.patch_size
            mask = bool_masked_pos.reshape(-1, size, size).repeat_interleave(cfg.patch_size, 1) \
                .repeat_interleave(cfg.patch_size, 2).unsqueeze(1).to(recon.dtype)
            l1 = (recon.float() - pixel_values.float()).abs() * mask
            res["loss"] = l1.sum() / (mask.sum() + 1e-5) / cfg.num_channels
        return res
