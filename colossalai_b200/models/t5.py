"""T5 family: `T5Model`, `T5ForConditionalGeneration`, `T5EncoderModel`, `T5ForTokenClassification`.

T5 specifics kept: RMS ("T5LayerNorm") pre-norm blocks without biases, un-scaled dot-product attention with a
bucketed relative-position bias that the first layer owns and every layer of the stack reuses (here the stack owns
one `relative_attention_bias` table and computes the `[1, heads, Sq, Sk]` bias once per forward), ReLU or gated-GELU
FFN, shared input/output embedding with the `d_model**-0.5` logit scale when tied.

Parity: reference `colossalai/shardformer/policies/t5.py:34-560` + `modeling/t5.py:27-800`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .encdec import (BlockConfig, DecoderBlock, EncoderBlock, init_module_weights, make_norm, seq2seq_loss,
                     shift_tokens_right)

__all__ = ["T5Config", "T5Stack", "T5Model", "T5ForConditionalGeneration", "T5EncoderModel",
           "T5ForTokenClassification", "T5_ZOO", "relative_position_bucket"]


@dataclass
class T5Config:
    model_type: str = "t5"
    vocab_size: int = 32128
    d_model: int = 512
    d_kv: int = 64
    d_ff: int = 2048
    num_layers: int = 6
    num_decoder_layers: Optional[int] = None
    num_heads: int = 8
    relative_attention_num_buckets: int = 32
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6
    feed_forward_proj: str = "relu"      # "relu" | "gated-gelu"
    tie_word_embeddings: bool = True
    dropout_rate: float = 0.0
    initializer_factor: float = 1.0
    pad_token_id: int = 0
    eos_token_id: int = 1
    decoder_start_token_id: int = 0
    num_labels: int = 2

    def __post_init__(self) -> None:
        if self.num_decoder_layers is None:
            self.num_decoder_layers = self.num_layers

    @property
    def hidden_size(self) -> int:
        return self.d_model

    def block(self) -> BlockConfig:
        gated = self.feed_forward_proj.startswith("gated")
        act = "gelu_new" if gated else self.feed_forward_proj
        return BlockConfig(hidden_size=self.d_model, num_heads=self.num_heads, head_dim=self.d_kv, ffn_dim=self.d_ff,
                           act=act, glu=gated, qkv_bias=False, out_bias=False, mlp_bias=False, norm_type="rms",
                           norm_eps=self.layer_norm_epsilon, pre_norm=True, attn_scale=1.0, dropout=self.dropout_rate)

    def replace(self, **kw) -> "T5Config":
        return replace(self, **kw)


T5_ZOO: Dict[str, T5Config] = {
    "t5-small": T5Config(),
    "t5-base": T5Config(d_model=768, d_ff=3072, num_layers=12, num_heads=12),
    "t5-large": T5Config(d_model=1024, d_ff=4096, num_layers=24, num_heads=16),
    "flan-t5-xl": T5Config(d_model=2048, d_ff=5120, num_layers=24, num_heads=32, feed_forward_proj="gated-gelu",
                           tie_word_embeddings=False),
    "t5-tiny": T5Config(vocab_size=512, d_model=64, d_kv=16, d_ff=128, num_layers=2, num_heads=4),
}


def relative_position_bucket(rel: torch.Tensor, bidirectional: bool, num_buckets: int, max_distance: int) -> torch.Tensor:
    """Map signed key-minus-query offsets to bucket ids: half exact small offsets, half log-spaced up to
    `max_distance` (separate halves for the two directions when bidirectional)."""
    ret = torch.zeros_like(rel)
    n = num_buckets
    if bidirectional:
        n //= 2
        ret = ret + (rel > 0).long() * n
        rel = rel.abs()
    else:
        rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = n // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float().clamp(min=1) / max_exact) / math.log(max_distance / max_exact)
                         * (n - max_exact)).long()
    large = torch.min(large, torch.full_like(large, n - 1))
    return ret + torch.where(is_small, rel, large)


class T5Stack(nn.Module):
    """One T5 stack over already-embedded inputs `[B, S, d_model]` (the owning model holds the shared embedding)."""

    def __init__(self, cfg: T5Config, is_decoder: bool) -> None:
        super().__init__()
        self.cfg, self.is_decoder = cfg, is_decoder
        bc = cfg.block()
        n = cfg.num_decoder_layers if is_decoder else cfg.num_layers
        self.block = nn.ModuleList([DecoderBlock(bc) if is_decoder else EncoderBlock(bc) for _ in range(n)])
        self.relative_attention_bias = nn.Embedding(cfg.relative_attention_num_buckets, cfg.num_heads)
        self.final_layer_norm = make_norm(bc)
        self.gradient_checkpointing = False

    def position_bias(self, q_len: int, k_len: int, device, offset: int = 0) -> torch.Tensor:
        ctx = torch.arange(offset, offset + q_len, device=device)[:, None]
        mem = torch.arange(k_len, device=device)[None, :]
        buckets = relative_position_bucket(mem - ctx, not self.is_decoder, self.cfg.relative_attention_num_buckets,
                                           self.cfg.relative_attention_max_distance)
        return self.relative_attention_bias(buckets).permute(2, 0, 1).unsqueeze(0)   # [1, heads_local, q, k]

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                memory: Optional[torch.Tensor] = None, memory_mask: Optional[torch.Tensor] = None,
                caches: Optional[List[dict]] = None) -> torch.Tensor:
        S = x.shape[1]
        past = 0
        if caches is not None and caches[0].get("self") is not None:
            past = caches[0]["self"][0].shape[2]
        bias = self.position_bias(S, past + S, x.device, offset=past)
        for i, blk in enumerate(self.block):
            if self.is_decoder:
                args = dict(memory=memory, self_bias=bias, key_padding_mask=attention_mask,
                            memory_padding_mask=memory_mask, cache=None if caches is None else caches[i])
                if self.gradient_checkpointing and self.training:
                    x = torch.utils.checkpoint.checkpoint(blk, x, use_reentrant=False, **args)
                else:
                    x = blk(x, **args)
            elif self.gradient_checkpointing and self.training:
                x = torch.utils.checkpoint.checkpoint(blk, x, bias, attention_mask, use_reentrant=False)
            else:
                x = blk(x, bias, attention_mask)
        return self.final_layer_norm(x)


class _T5Base(nn.Module):
    def __init__(self, cfg: T5Config) -> None:
        super().__init__()
        self.cfg = self.config = cfg
        self.shared = nn.Embedding(cfg.vocab_size, cfg.d_model)
        self.shard_config = None

    def _init(self) -> None:
        init_module_weights(self, 0.02 * self.cfg.initializer_factor)

    def gradient_checkpointing_enable(self, *a, **k) -> None:
        for m in self.modules():
            if isinstance(m, T5Stack):
                m.gradient_checkpointing = True

    def get_input_embeddings(self):
        return self.shared


class T5Model(_T5Base):
    def __init__(self, config: Optional[T5Config] = None, **kw) -> None:
        super().__init__(config or T5Config(**kw))
        self.encoder = T5Stack(self.cfg, False)
        self.decoder = T5Stack(self.cfg, True)
        self._init()

    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                decoder_input_ids: Optional[torch.Tensor] = None,
                decoder_attention_mask: Optional[torch.Tensor] = None,
                encoder_outputs: Optional[torch.Tensor] = None, caches=None, **unused) -> Dict[str, torch.Tensor]:
        mem = encoder_outputs if encoder_outputs is not None else self.encoder(self.shared(input_ids), attention_mask)
        h = self.decoder(self.shared(decoder_input_ids), decoder_attention_mask, memory=mem, memory_mask=attention_mask,
                         caches=caches)
        return {"last_hidden_state": h, "encoder_last_hidden_state": mem}


class T5ForConditionalGeneration(_T5Base):
    def __init__(self, config: Optional[T5Config] = None, **kw) -> None:
        super().__init__(config or T5Config(**kw))
        cfg = self.cfg
        self.encoder = T5Stack(cfg, False)
        self.decoder = T5Stack(cfg, True)
        self.lm_head = nn.Linear(cfg.d_model, cfg.vocab_size, bias=False)
        self._init()
        if cfg.tie_word_embeddings:
            self.lm_head.weight = self.shared.weight

    def get_output_embeddings(self):
        return self.lm_head

    def _logits(self, h: torch.Tensor) -> torch.Tensor:
        if self.cfg.tie_word_embeddings:
            h = h * (self.cfg.d_model ** -0.5)
        return self.lm_head(h)

    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                decoder_input_ids: Optional[torch.Tensor] = None,
                decoder_attention_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                encoder_outputs: Optional[torch.Tensor] = None, caches=None, **unused) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        if decoder_input_ids is None and labels is not None:
            decoder_input_ids = shift_tokens_right(labels, cfg.pad_token_id, cfg.decoder_start_token_id)
        mem = encoder_outputs if encoder_outputs is not None else self.encoder(self.shared(input_ids), attention_mask)
        h = self.decoder(self.shared(decoder_input_ids), decoder_attention_mask, memory=mem, memory_mask=attention_mask,
                         caches=caches)
        logits = self._logits(h)
        out = {"logits": logits, "encoder_last_hidden_state": mem}
        if labels is not None:
            out["loss"] = seq2seq_loss(logits, labels, self.shard_config, cfg.vocab_size)
        return out

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                 max_new_tokens: int = 20) -> torch.Tensor:
        """Greedy decoding with per-layer self/cross KV caches."""
        cfg = self.cfg
        mem = self.encoder(self.shared(input_ids), attention_mask)
        B = input_ids.shape[0]
        caches = [dict() for _ in self.decoder.block]
        cur = torch.full((B, 1), cfg.decoder_start_token_id, dtype=torch.long, device=input_ids.device)
        out = [cur]
        done = torch.zeros(B, dtype=torch.bool, device=input_ids.device)
        for _ in range(max_new_tokens):
            h = self.decoder(self.shared(cur), None, memory=mem, memory_mask=attention_mask, caches=caches)
            logits = self._logits(h[:, -1])
            if self.shard_config is not None and logits.shape[-1] != cfg.vocab_size and \
                    getattr(self.lm_head, "gather_output", True) is False:
                from ..parallel import comm

                logits = comm.all_gather(logits, -1, self.shard_config.tensor_parallel_process_group)
            cur = logits[..., :cfg.vocab_size].argmax(-1, keepdim=True)
            cur = cur.masked_fill(done[:, None], cfg.pad_token_id)
            out.append(cur)
            done |= cur.squeeze(1) == cfg.eos_token_id
            if bool(done.all()):
                break
        return torch.cat(out, dim=1)


class T5EncoderModel(_T5Base):
    def __init__(self, config: Optional[T5Config] = None, **kw) -> None:
        super().__init__(config or T5Config(**kw))
        self.encoder = T5Stack(self.cfg, False)
        self._init()

    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                **unused) -> Dict[str, torch.Tensor]:
        return {"last_hidden_state": self.encoder(self.shared(input_ids), attention_mask)}


class T5ForTokenClassification(_T5Base):
    def __init__(self, config: Optional[T5Config] = None, **kw) -> None:
        super().__init__(config or T5Config(**kw))
        self.encoder = T5Stack(self.cfg, False)
        self.classifier = nn.Linear(self.cfg.d_model, self.cfg.num_labels)
        self._init()

    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, **unused) -> Dict[str, torch.Tensor]:
        logits = self.classifier(self.encoder(self.shared(input_ids), attention_mask))
        out = {"logits": logits}
        if labels is not None:
            out["loss"] = nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(),
                                                      labels.reshape(-1), ignore_index=-100)
        return out
