"""Persistent split-KV workspace of the paged decode-attention kernel.

Parity: reference `colossalai/inference/flash_decoding_utils.py` (`FDIntermTensors`: `mid_output`, `mid_output_lse`,
`exp_sums`, `max_logits` allocated once for the largest batch).  `cb_paged_decode_attention` writes one partial
output and one (max, sum) pair per (sequence, head, split); keeping those buffers alive across layers and steps keeps
the allocator out of the decode loop and gives CUDA-graph capture fixed addresses."""
from __future__ import annotations

import torch

__all__ = ["FDIntermTensors"]


class _Singleton(type):
    _inst = {}

    def __call__(cls, *a, **k):
        if cls not in cls._inst:
            cls._inst[cls] = super().__call__(*a, **k)
        return cls._inst[cls]


class FDIntermTensors(metaclass=_Singleton):
    def __init__(self) -> None:
        self._tensors_initialized = False

    def _reset(self) -> None:
        self._tensors_initialized = False
        for n in ("_mid_output", "_mid_output_lse", "_exp_sums", "_max_logits"):
            if hasattr(self, n):
                delattr(self, n)

    @property
    def is_initialized(self) -> bool:
        return self._tensors_initialized

    @property
    def mid_output(self) -> torch.Tensor:
        """[max_batch, heads, splits, head_dim] fp32 partial outputs."""
        return self._mid_output

    @property
    def mid_output_lse(self) -> torch.Tensor:
        """[max_batch, heads, splits, 2] fp32 (running max, running sum) per split."""
        return self._mid_output_lse

    @property
    def exp_sums(self) -> torch.Tensor:
        return self._mid_output_lse[..., 1]

    @property
    def max_logits(self) -> torch.Tensor:
        return self._mid_output_lse[..., 0]

    def initialize(self, max_batch_size: int, num_attn_heads: int, kv_max_split_num: int, head_dim: int,
                   dtype: torch.dtype = torch.float32, device=None) -> None:
        assert not self._tensors_initialized, "workspace already initialised (call _reset() to re-size it)"
        device = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        self._mid_output = torch.empty(max_batch_size, num_attn_heads, kv_max_split_num, head_dim, dtype=dtype,
                                       device=device)
        self._mid_output_lse = torch.empty(max_batch_size, num_attn_heads, kv_max_split_num, 2, dtype=dtype,
                                           device=device)
        self._tensors_initialized = True

    def ensure(self, total_splits: int, num_attn_heads: int, head_dim: int, device=None) -> None:
        """Grow-only variant for several engines in one process (speculative decoding runs two): buffers that earlier
        CUDA graphs may still point at are retired, never freed."""
        need_o, need_l = total_splits * num_attn_heads * head_dim, total_splits * num_attn_heads * 2
        if self._tensors_initialized and self._mid_output.numel() >= need_o and self._mid_output_lse.numel() >= need_l \
                and (device is None or self._mid_output.device == torch.device(device)):
            return
        if self._tensors_initialized:
            self.__dict__.setdefault("_retired", []).append((self._mid_output, self._mid_output_lse))
            self._tensors_initialized = False
        self.initialize(1, num_attn_heads, total_splits, head_dim, dtype=torch.float32, device=device)

    def views(self, n_seqs: int, n_heads: int, splits: int, head_dim: int):
        """Contiguous `[n, H, splits, D]` / `[n, H, splits, 2]` windows for one call, or None when the workspace is
        too small (the op then allocates)."""
        if not self._tensors_initialized:
            return None
        need_o, need_l = n_seqs * n_heads * splits * head_dim, n_seqs * n_heads * splits * 2
        if need_o > self._mid_output.numel() or need_l > self._mid_output_lse.numel():
            return None
        o = self._mid_output.view(-1)[:need_o].view(n_seqs, n_heads, splits, head_dim)
        ml = self._mid_output_lse.view(-1)[:need_l].view(n_seqs, n_heads, splits, 2)
        return o, ml
