"""MoE kernels' python face (kernel/csrc/moe.cu) with PyTorch reference paths.

 * capacity-based dispatch/combine + cumsum (legacy MoE layer; reference `MoeDispatch`/`MoeCombine`/`moe_cumsum` in
   colossalai/moe/_operation.py:220-330 backed by moe_kernel.cu)
 * fused router softmax + top-k
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..kernel import loader
from ._dispatch import use_native
from ._dtypes import code

_lib = None


def _get_lib():
    global _lib
    if _lib is None:
        _lib = loader.load("cb200_moe")
    return _lib


def _i32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.int32 and t.is_contiguous() else t.to(torch.int32).contiguous()


def _expand_dest(mask: torch.Tensor, dest_idx: torch.Tensor) -> torch.Tensor:
    """Accept the reference's compact `dest_idx [s]` (one capacity slot per token) as well as the full `[s, e]` form."""
    if dest_idx.dim() == 1:
        return dest_idx[:, None].expand_as(mask)
    return dest_idx


def cumsum_sub_one(mask: torch.Tensor) -> torch.Tensor:
    """cumsum(mask, dim=0) - 1 for an integer [s, e] mask."""
    if use_native(mask) and mask.dim() == 2:
        m = _i32(mask)
        out = torch.empty_like(m)
        loader.check(_get_lib().cb_moe_cumsum_sub_one(loader.ptr(m), loader.ptr(out), m.shape[0], m.shape[1],
                                                      loader.stream_ptr()), "moe_cumsum")
        loader.launch_counter.add("moe_cumsum")
        return out.to(mask.dtype)
    return torch.cumsum(mask, dim=0) - 1


def dispatch_forward(s: int, ec: int, h: int, tokens: torch.Tensor, mask: torch.Tensor, dest_idx: torch.Tensor
                     ) -> torch.Tensor:
    """tokens [s, h] -> [e, c, h] where `ec` = e * c (total slots)."""
    e = mask.shape[1]
    c = ec // e
    dest = _expand_dest(mask, dest_idx)
    out = torch.zeros(e, c, h, dtype=tokens.dtype, device=tokens.device)
    if use_native(tokens):
        loader.check(_get_lib().cb_moe_dispatch_fwd(loader.ptr(tokens.contiguous()), loader.ptr(out),
                                                    loader.ptr(_i32(mask)), loader.ptr(_i32(dest)), s, e, c, h,
                                                    code(tokens.dtype), loader.stream_ptr()), "moe_dispatch_fwd")
        loader.launch_counter.add("moe_dispatch_fwd")
        return out
    tok_i, exp_i = torch.nonzero(mask, as_tuple=True)
    pos = dest[tok_i, exp_i].long()
    ok = (pos >= 0) & (pos < c)
    out[exp_i[ok], pos[ok]] = tokens[tok_i[ok]]
    return out


def dispatch_backward(s: int, ec: int, h: int, d_expert: torch.Tensor, mask: torch.Tensor, dest_idx: torch.Tensor
                      ) -> torch.Tensor:
    e = mask.shape[1]
    c = ec // e
    dest = _expand_dest(mask, dest_idx)
    if use_native(d_expert):
        out = torch.empty(s, h, dtype=d_expert.dtype, device=d_expert.device)
        loader.check(_get_lib().cb_moe_dispatch_bwd(loader.ptr(out), loader.ptr(d_expert.contiguous()),
                                                    loader.ptr(_i32(mask)), loader.ptr(_i32(dest)), s, e, c, h,
                                                    code(d_expert.dtype), loader.stream_ptr()), "moe_dispatch_bwd")
        loader.launch_counter.add("moe_dispatch_bwd")
        return out
    out = torch.zeros(s, h, dtype=torch.float32, device=d_expert.device)
    tok_i, exp_i = torch.nonzero(mask, as_tuple=True)
    pos = dest[tok_i, exp_i].long()
    ok = (pos >= 0) & (pos < c)
    out.index_add_(0, tok_i[ok], d_expert.view(e, c, h)[exp_i[ok], pos[ok]].float())
    return out.to(d_expert.dtype)


def combine_forward(s: int, e: int, c: int, h: int, expert_out: torch.Tensor, logits: torch.Tensor, mask: torch.Tensor,
                    dest_idx: torch.Tensor) -> torch.Tensor:
    dest = _expand_dest(mask, dest_idx)
    if use_native(expert_out):
        out = torch.empty(s, h, dtype=expert_out.dtype, device=expert_out.device)
        loader.check(_get_lib().cb_moe_combine_fwd(loader.ptr(expert_out.contiguous()), loader.ptr(out),
                                                   loader.ptr(logits.float().contiguous()), loader.ptr(_i32(mask)),
                                                   loader.ptr(_i32(dest)), s, e, c, h, code(expert_out.dtype),
                                                   loader.stream_ptr()), "moe_combine_fwd")
        loader.launch_counter.add("moe_combine_fwd")
        return out
    out = torch.zeros(s, h, dtype=torch.float32, device=expert_out.device)
    tok_i, exp_i = torch.nonzero(mask, as_tuple=True)
    pos = dest[tok_i, exp_i].long()
    ok = (pos >= 0) & (pos < c)
    rows = expert_out.view(e, c, h)[exp_i[ok], pos[ok]].float() * logits.float()[tok_i[ok], exp_i[ok]][:, None]
    out.index_add_(0, tok_i[ok], rows)
    return out.to(expert_out.dtype)


def combine_backward(s: int, e: int, c: int, h: int, dy: torch.Tensor, expert_out: torch.Tensor, logits: torch.Tensor,
                     mask: torch.Tensor, dest_idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    dest = _expand_dest(mask, dest_idx)
    if use_native(dy):
        d_expert = torch.zeros(e * c, h, dtype=dy.dtype, device=dy.device)
        d_logits = torch.empty(s, e, dtype=torch.float32, device=dy.device)
        loader.check(_get_lib().cb_moe_combine_bwd(loader.ptr(dy.contiguous()), loader.ptr(expert_out.contiguous()),
                                                   loader.ptr(d_expert), loader.ptr(d_logits),
                                                   loader.ptr(logits.float().contiguous()), loader.ptr(_i32(mask)),
                                                   loader.ptr(_i32(dest)), s, e, c, h, code(dy.dtype),
                                                   loader.stream_ptr()), "moe_combine_bwd")
        loader.launch_counter.add("moe_combine_bwd")
        return d_expert, d_logits.to(logits.dtype)
    d_expert = torch.zeros(e, c, h, dtype=torch.float32, device=dy.device)
    d_logits = torch.zeros(s, e, dtype=torch.float32, device=dy.device)
    tok_i, exp_i = torch.nonzero(mask, as_tuple=True)
    pos = dest[tok_i, exp_i].long()
    ok = (pos >= 0) & (pos < c)
    tok_i, exp_i, pos = tok_i[ok], exp_i[ok], pos[ok]
    g = dy.float()[tok_i]
    d_expert[exp_i, pos] = g * logits.float()[tok_i, exp_i][:, None]
    d_logits[tok_i, exp_i] = (g * expert_out.view(e, c, h)[exp_i, pos].float()).sum(-1)
    return d_expert.view(e * c, h).to(dy.dtype), d_logits.to(logits.dtype)


def router_topk(logits: torch.Tensor, k: int, renormalize: bool = True, return_probs: bool = False):
    """softmax over experts (fp32) then top-k; -> (weights fp32 [T,k], indices int64 [T,k][, probs fp32 [T,E]]).
    Forward-only helper (the training router keeps autograd through torch ops; inference and the fused EP path use this)."""
    T, E = logits.shape
    if use_native(logits) and E <= 256 and k <= 32 and not logits.requires_grad:
        w = torch.empty(T, k, dtype=torch.float32, device=logits.device)
        idx = torch.empty(T, k, dtype=torch.int32, device=logits.device)
        probs = torch.empty(T, E, dtype=torch.float32, device=logits.device) if return_probs else None
        loader.check(_get_lib().cb_moe_router_topk(loader.ptr(logits.contiguous()), loader.ptr(w), loader.ptr(idx),
                                                   loader.ptr(probs), T, E, k, int(renormalize), code(logits.dtype),
                                                   loader.stream_ptr()), "moe_router_topk")
        loader.launch_counter.add("moe_router_topk")
        return (w, idx.long(), probs) if return_probs else (w, idx.long())
    probs = torch.softmax(logits.float(), dim=-1)
    w, idx = torch.topk(probs, k, dim=-1)
    if renormalize:
        w = w / w.sum(-1, keepdim=True).clamp_min(1e-20)
    return (w, idx, probs) if return_probs else (w, idx)
