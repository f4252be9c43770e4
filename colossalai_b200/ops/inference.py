"""Python face of the paged-KV inference kernels (kernel/csrc/inference.cu) with PyTorch reference paths.
Parity: reference `inference_ops_cuda` bindings used by `inference/modeling/backends/*.py`."""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch

from ..kernel import loader
from ._dispatch import use_native
from ._dtypes import code

_lib = None


def _get_lib():
    global _lib
    if _lib is None:
        _lib = loader.load("cb200_infer")
    return _lib


def kv_cache_write(k: torch.Tensor, v: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                   block_tables: torch.Tensor, token_seq: torch.Tensor, token_pos: torch.Tensor) -> None:
    """k, v: [tokens, Hkv, D] -> caches [nb, bs, Hkv, D] at (block_tables[seq][pos // bs], pos % bs)."""
    tokens, Hkv, D = k.shape
    bs = k_cache.shape[1]
    if use_native(k) and k.dtype in (torch.float16, torch.bfloat16) and k_cache.dtype == k.dtype:
        lib = _get_lib()
        loader.check(lib.cb_kv_cache_write(loader.ptr(k), loader.ptr(v), loader.ptr(k_cache), loader.ptr(v_cache),
                                           loader.ptr(block_tables), loader.ptr(token_seq), loader.ptr(token_pos),
                                           tokens, Hkv, D, bs, block_tables.shape[1], ctypes.c_int64(k.stride(0)),
                                           ctypes.c_int64(v.stride(0)), code(k.dtype), loader.stream_ptr()),
                     "kv_cache_write")
        loader.launch_counter.add("kv_cache_write")
        return
    blk = block_tables[token_seq.long(), (token_pos // bs).long()].long()
    slot = (token_pos % bs).long()
    k_cache[blk, slot] = k.to(k_cache.dtype)
    v_cache[blk, slot] = v.to(v_cache.dtype)


def _decode_workspace(q: torch.Tensor, n: int, heads: int, splits: int, head_dim: int):
    """Persistent split-KV buffers (`inference.flash_decoding_utils.FDIntermTensors`) when the engine set them up."""
    from ..inference.flash_decoding_utils import FDIntermTensors

    fd = FDIntermTensors()
    if not fd.is_initialized or fd.mid_output.device != q.device or fd.mid_output.dtype != torch.float32:
        return None
    return fd.views(n, heads, splits, head_dim)


def paged_decode_attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, block_tables: torch.Tensor,
                           seq_lens: torch.Tensor, scale: Optional[float] = None,
                           alibi_slopes: Optional[torch.Tensor] = None, window: Optional[int] = None) -> torch.Tensor:
    """q: [num_seqs, Hq, D] (one new token per sequence) -> [num_seqs, Hq, D].  `window`: sliding-window attention
    (only the last `window` cached tokens are visible), handled inside the split-KV kernel (`window_start`)."""
    n, Hq, D = q.shape
    nb, bs, Hkv, _ = k_cache.shape
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    if (use_native(q) and q.dtype in (torch.float16, torch.bfloat16) and k_cache.dtype == q.dtype
            and D in (64, 128, 256) and Hq // Hkv <= 8):
        lib = _get_lib()
        max_len = block_tables.shape[1] * bs
        part = ctypes.c_int(0)
        splits = lib.cb_decode_num_splits(n, Hkv, max_len, ctypes.byref(part))
        out = torch.empty_like(q)
        ws = _decode_workspace(q, n, Hq, splits, D)
        if ws is not None:
            o_part, ml_part = ws
        else:
            o_part = torch.empty(n, Hq, splits, D, dtype=torch.float32, device=q.device)
            ml_part = torch.empty(n, Hq, splits, 2, dtype=torch.float32, device=q.device)
        qc = q if q.stride(2) == 1 and q.stride(1) == D else q.contiguous()
        loader.check(lib.cb_paged_decode_attention(
            loader.ptr(qc), loader.ptr(k_cache), loader.ptr(v_cache), loader.ptr(block_tables), loader.ptr(seq_lens),
            loader.ptr(out), loader.ptr(o_part), loader.ptr(ml_part), n, Hq, Hkv, D, bs, block_tables.shape[1], splits,
            part.value, ctypes.c_float(scale), loader.ptr(alibi_slopes), ctypes.c_int64(qc.stride(0)),
            ctypes.c_int64(out.stride(0)), code(q.dtype), int(window or 0), loader.stream_ptr()),
            "paged_decode_attention")
        loader.launch_counter.add("paged_decode_attention", 2)
        return out
    return paged_decode_attention_ref(q, k_cache, v_cache, block_tables, seq_lens, scale, alibi_slopes, window)


def paged_decode_attention_ref(q, k_cache, v_cache, block_tables, seq_lens, scale=None, alibi_slopes=None,
                               window=None):
    n, Hq, D = q.shape
    nb, bs, Hkv, _ = k_cache.shape
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    G = Hq // Hkv
    out = torch.empty_like(q)
    for i in range(n):
        L = int(seq_lens[i])
        nblk = (L + bs - 1) // bs
        blks = block_tables[i, :nblk].long()
        k = k_cache[blks].reshape(-1, Hkv, D)[:L].float()
        v = v_cache[blks].reshape(-1, Hkv, D)[:L].float()
        if window is not None and L > window:
            k, v, L = k[L - window:], v[L - window:], window
        k = k.repeat_interleave(G, dim=1)
        v = v.repeat_interleave(G, dim=1)
        s = torch.einsum("hd,lhd->hl", q[i].float(), k) * scale
        if alibi_slopes is not None:
            s = s + alibi_slopes.float()[:, None] * (torch.arange(L, device=q.device) - (L - 1))[None, :]
        p = s.softmax(-1)
        out[i] = torch.einsum("hl,lhd->hd", p, v).to(q.dtype)
    return out


def convert_fp8(x: torch.Tensor, to_fp8: bool, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """KV-cache fp8 (e5m2) storage conversion."""
    if to_fp8:
        if use_native(x):
            out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
            loader.check(_get_lib().cb_convert_fp8(loader.ptr(x.contiguous()), loader.ptr(out),
                                                   ctypes.c_int64(x.numel()), code(x.dtype), 0, loader.stream_ptr()),
                         "convert_fp8")
            loader.launch_counter.add("convert_fp8")
            return out
        return x.to(torch.float8_e5m2).view(torch.uint8)
    if use_native(x):
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        loader.check(_get_lib().cb_convert_fp8(loader.ptr(x.contiguous()), loader.ptr(out), ctypes.c_int64(x.numel()),
                                               code(out_dtype), 1, loader.stream_ptr()), "convert_fp8")
        loader.launch_counter.add("convert_fp8")
        return out
    return x.view(torch.float8_e5m2).to(out_dtype)
