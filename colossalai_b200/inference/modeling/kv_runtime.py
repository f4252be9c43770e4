"""Runtime object handed to the model as `kv_cache`: writes K/V into the paged cache and runs attention.

Replaces the reference's no-padding model forwards + attention backends (`inference/modeling/models/nopadding_llama.py`,
`modeling/backends/{attention_backend,pre_attention_backend}.py`): our generic model already works on flattened
un-padded tokens, so inference only has to (a) scatter the step's K/V into the block tables and (b) run varlen
flash-attention for prefill or the paged split-KV kernel for decode.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ... import ops
from ...ops import inference as iops

__all__ = ["PagedKVRuntime"]


class PagedKVRuntime:
    def __init__(self, k_caches: List[torch.Tensor], v_caches: List[torch.Tensor], block_size: int) -> None:
        self.k_caches, self.v_caches = k_caches, v_caches
        self.block_size = block_size
        self.block_tables: Optional[torch.Tensor] = None      # [bsz, max_blocks] int32 (device)
        self.seq_lens: Optional[torch.Tensor] = None          # [bsz] int32, INCLUDING the tokens of this step
        self.is_prompt = True
        self.token_seq: Optional[torch.Tensor] = None         # [tokens] int32 sequence index of every token
        self.token_pos: Optional[torch.Tensor] = None         # [tokens] int32 position of every token
        self.cu_seqlens: Optional[torch.Tensor] = None
        self.max_seqlen: int = 0
        self.q_per_seq: int = 1                               # decode: tokens verified per sequence (spec-dec)
        self.max_len_host: int = 0                            # longest sequence of the step (host side, no sync)

    def set_step(self, block_tables: torch.Tensor, seq_lens: torch.Tensor, is_prompt: bool, device,
                 q_per_seq: int = 1) -> torch.Tensor:
        """Prepare per-step metadata; returns the int64 position of every input token."""
        self.block_tables = block_tables.to(device=device, dtype=torch.int32).contiguous()
        self.seq_lens = seq_lens.to(device=device, dtype=torch.int32).contiguous()
        self.is_prompt = is_prompt
        self.q_per_seq = q_per_seq
        lens = seq_lens.tolist()
        self.max_len_host = max(lens) if lens else 0
        if is_prompt:
            seq = torch.repeat_interleave(torch.arange(len(lens)), torch.tensor(lens))
            pos = torch.cat([torch.arange(l) for l in lens])
            cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
            cu[1:] = torch.cumsum(torch.tensor(lens), 0)
            self.cu_seqlens = cu.to(device)
            self.max_seqlen = max(lens)
        else:
            seq = torch.repeat_interleave(torch.arange(len(lens)), q_per_seq)
            pos = torch.cat([torch.arange(l - q_per_seq, l) for l in lens])
            self.cu_seqlens = None
        self.token_seq = seq.to(device=device, dtype=torch.int32)
        self.token_pos = pos.to(device=device, dtype=torch.int32)
        return pos.to(device=device, dtype=torch.int64)

    def attend(self, layer_idx: int, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, meta, scale: float,
               alibi_slopes: Optional[torch.Tensor] = None, sliding_window: Optional[int] = None) -> torch.Tensor:
        kc, vc = self.k_caches[layer_idx], self.v_caches[layer_idx]
        iops.kv_cache_write(k, v, kc, vc, self.block_tables, self.token_seq, self.token_pos)
        if sliding_window is not None and self.max_len_host > sliding_window:
            return self._attend_windowed(q, k, v, kc, vc, scale, alibi_slopes, sliding_window)
        return self._attend_full(q, k, v, kc, vc, scale, alibi_slopes)

    def _native_prefill(self, q, k, v, scale, window: int, alibi_slopes) -> Optional[torch.Tensor]:
        """Sliding-window / ALiBi prefill inside the tcgen05 flash kernel (packed batch, device-resident boundaries);
        None when the tensors are outside the kernel's envelope (CPU tests, head_dim other than 64 / 128, fp32)."""
        from ...ops import flash_attn_native as fan

        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        if not fan.supported(q, k, v, self.cu_seqlens, self.cu_seqlens, force=True):
            return None
        return fan.flash_prefill(q, k, v, self.cu_seqlens, scale, window=window, alibi_slopes=alibi_slopes)

    def _attend_windowed(self, q, k, v, kc, vc, scale, alibi_slopes, window: int) -> torch.Tensor:
        """Sliding-window attention (Mistral): a sequence longer than the window takes the banded reference path —
        prefill with an explicit band mask per sequence, decode over the last `window` cached tokens."""
        from ...ops.attention import attention_ref

        if self.is_prompt:
            native = self._native_prefill(q, k, v, scale, window, alibi_slopes)
            if native is not None:
                return native
            cu = self.cu_seqlens.tolist()
            outs = []
            for i in range(len(cu) - 1):
                a, b = cu[i], cu[i + 1]
                S = b - a
                pos = torch.arange(S, device=q.device)
                rel = pos[None, :] - pos[:, None]
                keep = (rel <= 0) & (rel > -window)
                bias = torch.zeros(S, S, device=q.device, dtype=torch.float32)
                if alibi_slopes is not None:
                    bias = rel.clamp(max=0).float()[None] * alibi_slopes.float()[:, None, None]
                bias = bias.expand(q.shape[1], S, S) if bias.dim() == 3 else bias[None].expand(q.shape[1], S, S)
                bias = bias.masked_fill(~keep[None], float("-inf")).to(q.dtype)[None]
                outs.append(attention_ref(q[a:b], k[a:b], v[a:b], batch=1, causal=False, scale=scale, attn_mask=bias))
            return torch.cat(outs, 0)
        n = self.q_per_seq
        if n == 1:
            return iops.paged_decode_attention(q, kc, vc, self.block_tables, self.seq_lens, scale,
                                               alibi_slopes=alibi_slopes, window=window)
        bsz = self.seq_lens.numel()
        outs = []
        for j in range(n):
            qj = q.view(bsz, n, *q.shape[1:])[:, j]
            outs.append(iops.paged_decode_attention(qj.contiguous(), kc, vc, self.block_tables,
                                                    self.seq_lens - (n - 1 - j), scale, alibi_slopes=alibi_slopes,
                                                    window=window))
        return torch.stack(outs, 1).reshape(q.shape)

    def _attend_full(self, q, k, v, kc, vc, scale, alibi_slopes) -> torch.Tensor:
        if alibi_slopes is not None:
            # ALiBi families (Baichuan-13B, BLOOM): biased varlen prefill through the reference backend, the paged
            # decode kernel takes the slopes natively
            from .backends.attention_backend import AttentionMetaData, ReferenceAttentionBackend

            if self.is_prompt:
                native = self._native_prefill(q, k, v, scale, 0, alibi_slopes)
                if native is not None:
                    return native
                md = AttentionMetaData(q, k, v, kc, vc, self.block_tables, self.block_size, cu_seqlens=self.cu_seqlens,
                                       sm_scale=scale, alibi_slopes=alibi_slopes)
                return ReferenceAttentionBackend().prefill(md)
            assert self.q_per_seq == 1, "speculative verification is not implemented for ALiBi models"
            return iops.paged_decode_attention(q, kc, vc, self.block_tables, self.seq_lens, scale,
                                               alibi_slopes=alibi_slopes)
        if self.is_prompt:
            return ops.attention(q, k, v, causal=True, scale=scale, cu_seqlens_q=self.cu_seqlens,
                                 cu_seqlens_k=self.cu_seqlens, max_seqlen=self.max_seqlen)
        if self.q_per_seq == 1:
            return iops.paged_decode_attention(q, kc, vc, self.block_tables, self.seq_lens, scale)
        # speculative verification: n query tokens per sequence, each sees the cache up to its own position
        n = self.q_per_seq
        bsz = self.seq_lens.numel()
        outs = []
        for j in range(n):
            qj = q.view(bsz, n, *q.shape[1:])[:, j]
            outs.append(iops.paged_decode_attention(qj.contiguous(), kc, vc, self.block_tables,
                                                    self.seq_lens - (n - 1 - j), scale))
        return torch.stack(outs, 1).reshape(q.shape)
