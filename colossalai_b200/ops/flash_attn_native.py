"""ctypes face of the sm_100a flash-attention kernels (kernel/csrc/flash_attn_tcgen05.cu).  Filled in once the
kernel lands; until then `supported()` is False and `ops.attention` uses SDPA."""
from __future__ import annotations

import torch

_READY = False


def supported(q, k, v, cu_seqlens_q) -> bool:
    return _READY


def flash_attention(q, k, v, **kw):  # pragma: no cover
    raise NotImplementedError
