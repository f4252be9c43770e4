"""The original ring sequence parallelism: every rank keeps a `[B*heads, S/p, D]` slice of Q, K and V; the score
matrix `Q K^T` and the context `P V` are built by circulating the K (resp. V) sub-blocks around the ring.

Parity: reference `colossalai/legacy/nn/layer/parallel_sequence/_operation.py:15-160` (`RingQK`, `RingAV`) and
`layers.py:1-260` (`TransformerSelfAttentionRing`).  Communication uses `legacy.communication.ring_forward`
(one P2P hop per step over NVLink)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ....parallel import comm
from ...communication import ring_forward
from ...context import ParallelMode, global_context as gpc

__all__ = ["RingQK", "RingAV", "TransformerSelfAttentionRing"]


def _seq_group():
    return gpc.get_group(ParallelMode.SEQUENCE) if gpc.is_initialized(ParallelMode.SEQUENCE) else None


class RingQK(torch.autograd.Function):
    """scores[:, :, r*S_l:(r+1)*S_l] = Q_local @ K_r^T for every rank r, K blocks travelling around the ring."""

    @staticmethod
    def forward(ctx, sub_q: torch.Tensor, sub_k: torch.Tensor, batch_size: int, num_heads: int, sub_seq_len: int):
        g = _seq_group()
        p, r = comm.group_size(g), comm.group_rank(g)
        ctx.save_for_backward(sub_q, sub_k)
        ctx.sub_seq_len, ctx.p, ctx.r, ctx.g = sub_seq_len, p, r, g
        scores = sub_q.new_empty(batch_size * num_heads, sub_seq_len, sub_seq_len * p)
        k = sub_k
        for step in range(p):
            src = (r - step) % p          # after `step` hops we hold the block of rank r - step
            scores[:, :, src * sub_seq_len:(src + 1) * sub_seq_len] = torch.matmul(sub_q, k.transpose(1, 2))
            if step < p - 1:
                k = ring_forward(k, g)
        return scores

    @staticmethod
    def backward(ctx, grad: torch.Tensor):
        sub_q, sub_k = ctx.saved_tensors
        L, p, r, g = ctx.sub_seq_len, ctx.p, ctx.r, ctx.g
        # dK_r = sum over ranks of (their score-grad block r)^T @ their Q  -> reduce, keep own block
        grad_k = torch.matmul(grad.transpose(1, 2), sub_q)                     # [BH, L*p, D] partial for ALL blocks
        if p > 1:
            grad_k = comm.reduce_scatter(grad_k.contiguous(), 1, g)
        # dQ = sum_r grad[:, :, block r] @ K_r, K blocks circulate again
        grad_q = torch.zeros_like(sub_q)
        k = sub_k
        for step in range(p):
            src = (r - step) % p
            grad_q += torch.matmul(grad[:, :, src * L:(src + 1) * L], k)
            if step < p - 1:
                k = ring_forward(k, g)
        return grad_q, grad_k, None, None, None


class RingAV(torch.autograd.Function):
    """context = sum_r P[:, :, block r] @ V_r with V blocks travelling around the ring."""

    @staticmethod
    def forward(ctx, attention_score: torch.Tensor, sub_v: torch.Tensor, batch_size: int, num_heads: int,
                attention_head_size: int, sub_seq_len: int):
        g = _seq_group()
        p, r = comm.group_size(g), comm.group_rank(g)
        ctx.save_for_backward(attention_score, sub_v)
        ctx.sub_seq_len, ctx.p, ctx.r, ctx.g = sub_seq_len, p, r, g
        out = attention_score.new_zeros(batch_size * num_heads, sub_seq_len, attention_head_size)
        v = sub_v
        for step in range(p):
            src = (r - step) % p
            out += torch.matmul(attention_score[:, :, src * sub_seq_len:(src + 1) * sub_seq_len], v)
            if step < p - 1:
                v = ring_forward(v, g)
        return out

    @staticmethod
    def backward(ctx, grad: torch.Tensor):
        score, sub_v = ctx.saved_tensors
        L, p, r, g = ctx.sub_seq_len, ctx.p, ctx.r, ctx.g
        grad_v = torch.matmul(score.transpose(1, 2), grad)                       # partial for all V blocks
        if p > 1:
            grad_v = comm.reduce_scatter(grad_v.contiguous(), 1, g)
        grad_score = torch.zeros_like(score)
        v = sub_v
        for step in range(p):
            src = (r - step) % p
            grad_score[:, :, src * L:(src + 1) * L] = torch.matmul(grad, v.transpose(1, 2))
            if step < p - 1:
                v = ring_forward(v, g)
        return grad_score, grad_v, None, None, None, None


class TransformerSelfAttentionRing(nn.Module):
    """Self-attention over a sequence-sharded `[S/p, B, H]` input (Megatron layout) using RingQK / RingAV."""

    def __init__(self, hidden_size: int, num_attention_heads: int, attention_dropout: float = 0.0,
                 causal: bool = False) -> None:
        super().__init__()
        assert hidden_size % num_attention_heads == 0
        self.hidden_size, self.num_heads = hidden_size, num_attention_heads
        self.head_dim = hidden_size // num_attention_heads
        self.query_key_value = nn.Linear(hidden_size, 3 * hidden_size)
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.attention_dropout = attention_dropout
        self.causal = causal

    def forward(self, hidden_states: torch.Tensor, attention_mask: torch.Tensor = None) -> torch.Tensor:
        Sl, B, H = hidden_states.shape
        g = _seq_group()
        p, r = comm.group_size(g), comm.group_rank(g)
        qkv = self.query_key_value(hidden_states).view(Sl, B * self.num_heads, 3 * self.head_dim).transpose(0, 1)
        q, k, v = qkv.chunk(3, dim=-1)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        scores = RingQK.apply(q, k, B, self.num_heads, Sl) / math.sqrt(self.head_dim)
        if self.causal:
            qpos = torch.arange(r * Sl, (r + 1) * Sl, device=scores.device)[:, None]
            kpos = torch.arange(Sl * p, device=scores.device)[None, :]
            scores = scores.masked_fill(kpos > qpos, float("-inf"))
        if attention_mask is not None:          # additive [B, 1, 1|Sl, S] mask
            scores = (scores.view(B, self.num_heads, Sl, Sl * p) + attention_mask).view(B * self.num_heads, Sl, Sl * p)
        probs = F.softmax(scores.float(), dim=-1).to(scores.dtype)
        if self.attention_dropout > 0 and self.training:
            probs = F.dropout(probs, self.attention_dropout)
        ctxt = RingAV.apply(probs, v, B, self.num_heads, self.head_dim, Sl)      # [B*h, Sl, D]
        ctxt = ctxt.transpose(0, 1).reshape(Sl, B, H)
        return self.dense(ctxt)
