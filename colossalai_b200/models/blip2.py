"""BLIP-2: frozen-style ViT image encoder -> Q-Former (learned query tokens cross-attending to the image) ->
linear projection -> a causal language model that consumes the projected queries as a soft prefix.

The language model is our generic decoder (`TransformerLMHeadModel`, OPT configuration by default), fed through its
`inputs_embeds` path, so every tensor/sequence-parallel feature of the decoder stack applies to it unchanged.

Parity: reference `colossalai/shardformer/policies/blip2.py:20-420` + `modeling/blip2.py:14-120`
(`Blip2Model`, `Blip2ForConditionalGeneration`).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Dict, Optional

import torch
import torch.nn as nn

from .config import ModelConfig
from .encdec import BlockConfig, DecoderBlock, EncoderBlock, init_module_weights, make_norm
from .transformer import TransformerLMHeadModel

__all__ = ["Blip2Config", "Blip2VisionModel", "Blip2QFormer", "Blip2Model", "Blip2ForConditionalGeneration",
           "BLIP2_ZOO"]


def _opt_cfg(**kw) -> ModelConfig:
    base = dict(model_type="opt", vocab_size=50272, hidden_size=2560, intermediate_size=10240, num_hidden_layers=32,
                num_attention_heads=32, max_position_embeddings=2048, norm_type="layer", hidden_act="relu", glu=False,
                attention_bias=True, mlp_bias=True, pos_type="learned", tie_word_embeddings=True)
    base.update(kw)
    return ModelConfig(**base)


@dataclass
class Blip2Config:
    model_type: str = "blip2"
    # vision tower (EVA-CLIP-g shape by default)
    image_size: int = 224
    patch_size: int = 14
    vision_hidden_size: int = 1408
    vision_layers: int = 39
    vision_heads: int = 16
    vision_intermediate_size: int = 6144
    vision_eps: float = 1e-6
    # Q-Former
    num_query_tokens: int = 32
    qformer_hidden_size: int = 768
    qformer_layers: int = 12
    qformer_heads: int = 12
    qformer_intermediate_size: int = 3072
    cross_attention_frequency: int = 2
    qformer_eps: float = 1e-12
    # language model
    text_config: ModelConfig = field(default_factory=_opt_cfg)
    initializer_range: float = 0.02

    @property
    def hidden_size(self) -> int:
        return self.text_config.hidden_size

    @property
    def vocab_size(self) -> int:
        return self.text_config.vocab_size

    def vision_block(self) -> BlockConfig:
        return BlockConfig(hidden_size=self.vision_hidden_size, num_heads=self.vision_heads,
                           ffn_dim=self.vision_intermediate_size, act="gelu", norm_eps=self.vision_eps, pre_norm=True)

    def qformer_block(self) -> BlockConfig:
        return BlockConfig(hidden_size=self.qformer_hidden_size, num_heads=self.qformer_heads,
                           ffn_dim=self.qformer_intermediate_size, act="gelu", norm_eps=self.qformer_eps,
                           pre_norm=False, kv_hidden_size=self.vision_hidden_size)

    def replace(self, **kw) -> "Blip2Config":
        return replace(self, **kw)


BLIP2_ZOO: Dict[str, Blip2Config] = {
    "blip2-opt-2.7b": Blip2Config(),
    "blip2-tiny": Blip2Config(image_size=32, patch_size=8, vision_hidden_size=64, vision_layers=2, vision_heads=4,
                              vision_intermediate_size=128, num_query_tokens=4, qformer_hidden_size=64,
                              qformer_layers=2, qformer_heads=4, qformer_intermediate_size=128,
                              text_config=_opt_cfg(vocab_size=512, hidden_size=64, intermediate_size=128,
                                                   num_hidden_layers=2, num_attention_heads=4,
                                                   max_position_embeddings=128)),
}


class Blip2VisionModel(nn.Module):
    def __init__(self, cfg: Blip2Config) -> None:
        super().__init__()
        self.cfg = cfg
        H = cfg.vision_hidden_size
        n = (cfg.image_size // cfg.patch_size) ** 2
        self.patch_embedding = nn.Conv2d(3, H, cfg.patch_size, stride=cfg.patch_size)
        self.class_embedding = nn.Parameter(torch.zeros(1, 1, H))
        self.position_embedding = nn.Parameter(torch.zeros(1, n + 1, H))
        bc = cfg.vision_block()
        self.layers = nn.ModuleList([EncoderBlock(bc) for _ in range(cfg.vision_layers)])
        self.post_layernorm = make_norm(bc)
        self.gradient_checkpointing = False

    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        x = self.patch_embedding(pixel_values.to(self.patch_embedding.weight.dtype)).flatten(2).transpose(1, 2)
        x = torch.cat([self.class_embedding.expand(x.shape[0], -1, -1).to(x.dtype), x], dim=1)
        x = x + self.position_embedding[:, : x.shape[1]].to(x.dtype)
        for blk in self.layers:
            if self.gradient_checkpointing and self.training:
                x = torch.utils.checkpoint.checkpoint(blk, x, use_reentrant=False)
            else:
                x = blk(x)
        return self.post_layernorm(x)


class Blip2QFormer(nn.Module):
    """BERT-style post-norm blocks over the query tokens; every `cross_attention_frequency`-th block also
    cross-attends to the image features."""

    def __init__(self, cfg: Blip2Config) -> None:
        super().__init__()
        self.cfg = cfg
        bc = cfg.qformer_block()
        self.layernorm = make_norm(bc)
        self.layers = nn.ModuleList([
            DecoderBlock(bc, cross=(i % cfg.cross_attention_frequency == 0)) for i in range(cfg.qformer_layers)])
        for blk in self.layers:
            blk.self_attn.causal = False      # queries see each other bidirectionally

    def forward(self, query_embeds: torch.Tensor, image_embeds: torch.Tensor) -> torch.Tensor:
        x = self.layernorm(query_embeds)
        for blk in self.layers:
            x = blk(x, memory=image_embeds if hasattr(blk, "cross_attn") else None)
        return x


class Blip2Model(nn.Module):
    def __init__(self, config: Optional[Blip2Config] = None, **kw) -> None:
        super().__init__()
        cfg = config or Blip2Config(**kw)
        self.cfg = self.config = cfg
        self.vision_model = Blip2VisionModel(cfg)
        self.query_tokens = nn.Parameter(torch.zeros(1, cfg.num_query_tokens, cfg.qformer_hidden_size))
        self.qformer = Blip2QFormer(cfg)
        self.language_projection = nn.Linear(cfg.qformer_hidden_size, cfg.text_config.hidden_size)
        init_module_weights(self, cfg.initializer_range)
        with torch.no_grad():
            if self.query_tokens.device.type != "meta":
                self.query_tokens.normal_(0.0, cfg.initializer_range)
        self.language_model = TransformerLMHeadModel(cfg.text_config)
        self.shard_config = None

    def gradient_checkpointing_enable(self, *a, **k) -> None:
        self.vision_model.gradient_checkpointing = True
        self.language_model.gradient_checkpointing_enable()

    def get_image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        return self.vision_model(pixel_values)

    def get_qformer_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        img = self.vision_model(pixel_values)
        return self.qformer(self.query_tokens.expand(img.shape[0], -1, -1).to(img.dtype), img)

    def _prefix_and_text(self, pixel_values, input_ids):
        q = self.language_projection(self.get_qformer_features(pixel_values))
        txt = self.language_model.model.embed_tokens(input_ids)
        return torch.cat([q, txt.to(q.dtype)], dim=1), q.shape[1]

    def forward(self, pixel_values: torch.Tensor, input_ids: torch.Tensor,
                attention_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                **unused) -> Dict[str, torch.Tensor]:
        embeds, nq = self._prefix_and_text(pixel_values, input_ids)
        B = input_ids.shape[0]
        full_labels = None
        if labels is not None:
            full_labels = torch.cat([labels.new_full((B, nq), -100), labels], dim=1)
        out = self.language_model(inputs_embeds=embeds, labels=full_labels)
        if "logits" in out:
            lg = out["logits"]
            out["logits"] = lg.reshape(B, -1, lg.shape[-1])
        return out


class Blip2ForConditionalGeneration(Blip2Model):
    @torch.no_grad()
    def generate(self, pixel_values: torch.Tensor, input_ids: Optional[torch.Tensor] = None,
                 max_new_tokens: int = 20) -> torch.Tensor:
        """Greedy captioning (recomputes the prefix each step — the engine-grade KV-cached path is
        `colossalai_b200.inference`)."""
        tc = self.cfg.text_config
        B = pixel_values.shape[0]
        if input_ids is None:
            input_ids = torch.full((B, 1), tc.bos_token_id, dtype=torch.long, device=pixel_values.device)
        q = self.language_projection(self.get_qformer_features(pixel_values))
        ids = input_ids
        for _ in range(max_new_tokens):
            emb = torch.cat([q, self.language_model.model.embed_tokens(ids).to(q.dtype)], dim=1)
            lg = self.language_model(inputs_embeds=emb)["logits"]
            lg = lg.reshape(B, -1, lg.shape[-1])[:, -1, : tc.vocab_size]
            nxt = lg.argmax(-1, keepdim=True)
            ids = torch.cat([ids, nxt], dim=1)
            if bool((nxt == tc.eos_token_id).all()):
                break
        return ids
