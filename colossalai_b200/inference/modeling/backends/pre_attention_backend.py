"""Pre-attention step of the inference forward: rotary embedding of q/k and the scatter of this step's K/V into the
paged cache.  Parity: reference `colossalai/inference/modeling/backends/pre_attention_backend.py:19-146`
(`CudaPreAttentionBackend`, `TritonPreAttentionBackend`, `get_pre_attention_backend`)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import torch

from .... import ops
from ....ops import inference as iops
from .attention_backend import AttentionMetaData

__all__ = ["PreAttentionBackend", "CudaPreAttentionBackend", "ReferencePreAttentionBackend",
           "get_pre_attention_backend"]


class PreAttentionBackend(ABC):
    @abstractmethod
    def prefill(self, attn_metadata: AttentionMetaData, **kwargs) -> None:
        ...

    @abstractmethod
    def decode(self, attn_metadata: AttentionMetaData, **kwargs) -> None:
        ...


class CudaPreAttentionBackend(PreAttentionBackend):
    """RoPE through the fused qkv rotary kernel, then one kernel scatters K/V rows into their cache blocks."""

    def __init__(self, use_alibi_attn: bool = False) -> None:
        self.use_alibi_attn = use_alibi_attn

    def _run(self, m: AttentionMetaData, cos=None, sin=None, positions=None, token_seq=None, token_pos=None,
             rot_dim: Optional[int] = None, interleaved: bool = False) -> None:
        q, k, v = m.query_states, m.key_states, m.value_states
        if not self.use_alibi_attn and cos is not None:
            T, hq, D = q.shape
            hkv = k.shape[1]
            qkv = torch.cat([q.reshape(T, -1), k.reshape(T, -1), v.reshape(T, -1)], dim=-1)
            qkv = ops.rope_qkv(qkv, positions, cos, sin, hq, hkv, D, rot_dim=rot_dim or D, interleaved=interleaved)
            q2, k2, _ = qkv.split([hq * D, hkv * D, hkv * D], dim=-1)
            m.query_states, m.key_states = q2.reshape(T, hq, D), k2.reshape(T, hkv, D)
        iops.kv_cache_write(m.key_states, m.value_states, m.k_cache, m.v_cache, m.block_tables, token_seq, token_pos)

    def prefill(self, m: AttentionMetaData, **kwargs) -> None:
        self._run(m, **kwargs)

    def decode(self, m: AttentionMetaData, **kwargs) -> None:
        self._run(m, **kwargs)


class ReferencePreAttentionBackend(CudaPreAttentionBackend):
    """Same contract on any device (`ops.rope_qkv` / `kv_cache_write` fall back to PyTorch off-GPU)."""


def get_pre_attention_backend(model_shard_infer_config=None, use_alibi_attn: bool = False) -> PreAttentionBackend:
    use_cuda = getattr(model_shard_infer_config, "use_cuda_kernel", True)
    if use_cuda and torch.cuda.is_available():
        return CudaPreAttentionBackend(use_alibi_attn)
    return ReferencePreAttentionBackend(use_alibi_attn)
