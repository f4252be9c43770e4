"""Building blocks for the non-decoder-only families (ViT, T5, Whisper, BLIP-2, SAM): batch-major `[B, S, H]`
multi-head attention (self / cross, optional additive position bias), feed-forward, encoder and decoder blocks.

The blocks are ordinary single-device modules built from `nn.Linear`; the family policies swap the projections for
column / row parallel linears (self-attention uses ONE fused QKV GEMM, cross-attention a Q GEMM plus a fused KV GEMM)
and divide `num_heads` by the tensor-parallel size — the forward derives every shape from the local head count.

Parity: the per-family attention / layer forwards the reference patches in
`colossalai/shardformer/modeling/{vit.py:18-390, t5.py:27-800, whisper.py:36-1000, blip2.py:14-120, sam.py:8-220}`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..shardformer.layer.normalization import FusedLayerNorm, FusedRMSNorm

__all__ = ["BlockConfig", "MultiHeadAttention", "FeedForward", "EncoderBlock", "DecoderBlock", "make_norm",
           "init_module_weights", "shift_tokens_right", "seq2seq_loss"]


@dataclass
class BlockConfig:
    hidden_size: int
    num_heads: int
    ffn_dim: int
    head_dim: Optional[int] = None
    act: str = "gelu"
    glu: bool = False
    qkv_bias: bool = True
    out_bias: bool = True
    mlp_bias: bool = True
    norm_type: str = "layer"           # "layer" | "rms"
    norm_eps: float = 1e-5
    pre_norm: bool = True
    attn_scale: Optional[float] = None  # None -> 1/sqrt(head_dim); T5 uses 1.0
    dropout: float = 0.0
    kv_hidden_size: Optional[int] = None  # cross-attention memory width (defaults to hidden_size)

    def __post_init__(self) -> None:
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_heads


def make_norm(cfg: BlockConfig, hidden: Optional[int] = None) -> nn.Module:
    h = hidden or cfg.hidden_size
    return FusedRMSNorm(h, eps=cfg.norm_eps) if cfg.norm_type == "rms" else FusedLayerNorm(h, eps=cfg.norm_eps)


class MultiHeadAttention(nn.Module):
    """Self-attention (`cross=False`: fused `qkv_proj`) or cross-attention (`cross=True`: `q_proj` + fused
    `kv_proj` over the encoder memory).  `bias` is an additive `[1|B, heads_local, Sq, Sk]` term (T5 relative
    positions, SAM decomposed relative positions); `key_padding_mask` is `[B, Sk]` with 1 = keep."""

    def __init__(self, cfg: BlockConfig, cross: bool = False, causal: bool = False) -> None:
        super().__init__()
        self.cfg, self.cross, self.causal = cfg, cross, causal
        self.num_heads, self.head_dim = cfg.num_heads, cfg.head_dim
        inner = cfg.num_heads * cfg.head_dim
        if cross:
            self.q_proj = nn.Linear(cfg.hidden_size, inner, bias=cfg.qkv_bias)
            self.kv_proj = nn.Linear(cfg.kv_hidden_size or cfg.hidden_size, 2 * inner, bias=cfg.qkv_bias)
        else:
            self.qkv_proj = nn.Linear(cfg.hidden_size, 3 * inner, bias=cfg.qkv_bias)
        self.o_proj = nn.Linear(inner, cfg.hidden_size, bias=cfg.out_bias)
        self.scale = cfg.attn_scale if cfg.attn_scale is not None else 1.0 / math.sqrt(cfg.head_dim)
        self.shard_config = None
        self.kv_exchange = None

    def project_memory(self, memory: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """K/V of the encoder memory `[B, Sk, H]` -> two `[B, heads_local, Sk, D]` tensors (cacheable across decode steps)."""
        B, Sk, _ = memory.shape
        kv = self.kv_proj(memory)
        h = kv.shape[-1] // (2 * self.head_dim)
        k, v = kv.split(h * self.head_dim, dim=-1)
        return (k.reshape(B, Sk, h, self.head_dim).transpose(1, 2), v.reshape(B, Sk, h, self.head_dim).transpose(1, 2))

    def forward(self, x: torch.Tensor, memory: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                key_padding_mask: Optional[torch.Tensor] = None,
                memory_kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                past_kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, return_kv: bool = False):
        B, Sq, _ = x.shape
        D = self.head_dim
        if self.cross:
            q = self.q_proj(x)
            h = q.shape[-1] // D
            q = q.reshape(B, Sq, h, D).transpose(1, 2)
            k, v = memory_kv if memory_kv is not None else self.project_memory(memory)
        else:
            qkv = self.qkv_proj(x)
            h = qkv.shape[-1] // (3 * D)
            q, k, v = qkv.split(h * D, dim=-1)
            q, k, v = (t.reshape(B, Sq, h, D).transpose(1, 2) for t in (q, k, v))
            if past_kv is not None:
                k, v = torch.cat([past_kv[0], k], dim=2), torch.cat([past_kv[1], v], dim=2)
            if self.kv_exchange is not None:      # patch parallelism: keys/values of the other ranks' tokens
                k, v = self.kv_exchange(k, v, token_dim=2)
        Sk = k.shape[2]
        mask = bias
        causal = self.causal and not self.cross and Sq > 1
        if key_padding_mask is not None or (causal and (mask is not None or Sq != Sk)):
            keep = torch.ones(B, 1, Sq, Sk, dtype=torch.bool, device=x.device)
            if causal:
                keep = keep & torch.ones(Sq, Sk, dtype=torch.bool, device=x.device).tril(diagonal=Sk - Sq)
            if key_padding_mask is not None:
                keep = keep & key_padding_mask.bool()[:, None, None, :]
            add = torch.zeros(B, 1, Sq, Sk, dtype=q.dtype, device=x.device).masked_fill(~keep, float("-inf"))
            mask = add if mask is None else mask.to(q.dtype) + add
            causal = False
        elif mask is not None:
            mask = mask.to(q.dtype)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, is_causal=causal and mask is None,
                                           dropout_p=self.cfg.dropout if self.training else 0.0, scale=self.scale)
        o = self.o_proj(o.transpose(1, 2).reshape(B, Sq, h * D))
        return (o, (k, v)) if return_kv else o


class FeedForward(nn.Module):
    def __init__(self, cfg: BlockConfig) -> None:
        super().__init__()
        self.cfg = cfg
        if cfg.glu:
            self.gate_up_proj = nn.Linear(cfg.hidden_size, 2 * cfg.ffn_dim, bias=cfg.mlp_bias)
        else:
            self.up_proj = nn.Linear(cfg.hidden_size, cfg.ffn_dim, bias=cfg.mlp_bias)
        self.down_proj = nn.Linear(cfg.ffn_dim, cfg.hidden_size, bias=cfg.mlp_bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.cfg.glu:
            h = ops.glu(self.gate_up_proj(x), self.cfg.act)
        else:
            h = ops.get_activation(self.cfg.act)(self.up_proj(x))
        if self.cfg.dropout > 0 and self.training:
            h = F.dropout(h, self.cfg.dropout)
        return self.down_proj(h)


class EncoderBlock(nn.Module):
    """norm -> self-attention -> residual, norm -> FFN -> residual (pre-norm) or the post-norm (BERT) order."""

    def __init__(self, cfg: BlockConfig, causal: bool = False) -> None:
        super().__init__()
        self.cfg = cfg
        self.norm1 = make_norm(cfg)
        self.self_attn = MultiHeadAttention(cfg, cross=False, causal=causal)
        self.norm2 = make_norm(cfg)
        self.mlp = FeedForward(cfg)

    def forward(self, x: torch.Tensor, bias: Optional[torch.Tensor] = None,
                key_padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.cfg.pre_norm:
            x = x + self.self_attn(self.norm1(x), bias=bias, key_padding_mask=key_padding_mask)
            return x + self.mlp(self.norm2(x))
        x = self.norm1(x + self.self_attn(x, bias=bias, key_padding_mask=key_padding_mask))
        return self.norm2(x + self.mlp(x))


class DecoderBlock(nn.Module):
    """Causal self-attention, cross-attention over the encoder memory, FFN."""

    def __init__(self, cfg: BlockConfig, cross: bool = True) -> None:
        super().__init__()
        self.cfg = cfg
        self.norm1 = make_norm(cfg)
        self.self_attn = MultiHeadAttention(cfg, cross=False, causal=True)
        if cross:
            self.norm_cross = make_norm(cfg)
            self.cross_attn = MultiHeadAttention(cfg, cross=True)
        self.norm2 = make_norm(cfg)
        self.mlp = FeedForward(cfg)

    def forward(self, x: torch.Tensor, memory: Optional[torch.Tensor] = None, self_bias: Optional[torch.Tensor] = None,
                key_padding_mask: Optional[torch.Tensor] = None, memory_padding_mask: Optional[torch.Tensor] = None,
                cache: Optional[dict] = None) -> torch.Tensor:
        past = cache.get("self") if cache is not None else None
        mem_kv = cache.get("cross") if cache is not None else None
        has_cross = hasattr(self, "cross_attn") and (memory is not None or mem_kv is not None)
        if has_cross and cache is not None and mem_kv is None:
            mem_kv = cache["cross"] = self.cross_attn.project_memory(memory)

        def sa(h):
            if cache is None:
                return self.self_attn(h, bias=self_bias, key_padding_mask=key_padding_mask)
            o, kv = self.self_attn(h, bias=self_bias, key_padding_mask=key_padding_mask, past_kv=past, return_kv=True)
            cache["self"] = kv
            return o

        def ca(h):
            return self.cross_attn(h, memory=memory, key_padding_mask=memory_padding_mask, memory_kv=mem_kv)

        if self.cfg.pre_norm:
            x = x + sa(self.norm1(x))
            if has_cross:
                x = x + ca(self.norm_cross(x))
            return x + self.mlp(self.norm2(x))
        x = self.norm1(x + sa(x))
        if has_cross:
            x = self.norm_cross(x + ca(x))
        return self.norm2(x + self.mlp(x))


def init_module_weights(module: nn.Module, std: float = 0.02) -> None:
    """Normal(0, std) for linears / embeddings / convs, ones/zeros for norms; skips meta (lazy) parameters."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv1d, nn.Conv2d, nn.ConvTranspose2d)):
            if m.weight.device.type != "meta":
                nn.init.normal_(m.weight, mean=0.0, std=std)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Embedding):
            if m.weight.device.type != "meta":
                nn.init.normal_(m.weight, mean=0.0, std=std)


def shift_tokens_right(labels: torch.Tensor, pad_token_id: int, decoder_start_token_id: int) -> torch.Tensor:
    """Teacher forcing: decoder inputs are the labels shifted right behind the start token; -100 -> pad."""
    shifted = labels.new_zeros(labels.shape)
    shifted[:, 1:] = labels[:, :-1]
    shifted[:, 0] = decoder_start_token_id
    return shifted.masked_fill(shifted == -100, pad_token_id)


class _NoShard:
    enable_tensor_parallelism = False
    enable_sequence_parallelism = False
    parallel_output = False
    tensor_parallel_process_group = None
    sequence_parallel_process_group = None
    sequence_parallelism_mode = None


def seq2seq_loss(logits: torch.Tensor, labels: torch.Tensor, shard_config, vocab_size: int) -> torch.Tensor:
    """Un-shifted token cross entropy over (possibly vocab-parallel) logits `[B, S, V(/tp)]`."""
    from ..shardformer.layer.loss import dist_cross_entropy

    return dist_cross_entropy(labels, logits, shard_config if shard_config is not None else _NoShard, vocab_size,
                              dtype=torch.float32, shift=False)
