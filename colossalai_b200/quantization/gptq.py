"""GPTQ post-training weight quantisation (4 / 8 bit, per-group scales): the second-order, column-by-column
error-compensating solver, the packed `QuantLinear` it produces and the model-level driver.

Parity: reference `colossalai/legacy/inference/quant/gptq/` (`cai_gptq/cai_quant_linear.py:1-360` `CaiQuantLinear`
with packed `qweight` / `qzeros` / `scales` / `g_idx`, `gptq_manager.py`, built on auto-gptq's solver).  The solver
below is written from the algorithm: H = 2 X X^T, damped Cholesky inverse, blocked lazy updates of the not yet
quantised columns.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch
import torch.nn as nn

__all__ = ["Quantizer", "GPTQ", "QuantLinear", "CaiQuantLinear", "gptq_quantize_model", "pack_rows", "unpack_rows"]


class Quantizer:
    """Asymmetric min/max uniform quantiser; parameters are per output row (or per row x group)."""

    def __init__(self, bits: int = 4, sym: bool = False) -> None:
        self.bits, self.sym = bits, sym
        self.maxq = 2 ** bits - 1
        self.scale: Optional[torch.Tensor] = None
        self.zero: Optional[torch.Tensor] = None

    def find_params(self, w: torch.Tensor) -> None:
        """`w` [rows, cols] -> scale / zero [rows, 1]."""
        lo = w.min(dim=1).values.clamp(max=0)
        hi = w.max(dim=1).values.clamp(min=0)
        if self.sym:
            hi = torch.maximum(lo.abs(), hi)
            lo = -hi
        flat = (lo == 0) & (hi == 0)
        lo = torch.where(flat, -torch.ones_like(lo), lo)
        hi = torch.where(flat, torch.ones_like(hi), hi)
        self.scale = ((hi - lo) / self.maxq).unsqueeze(1)
        self.zero = (torch.full_like(lo, (self.maxq + 1) / 2) if self.sym else torch.round(-lo / self.scale[:, 0])).unsqueeze(1)

    def quantize_int(self, w: torch.Tensor) -> torch.Tensor:
        return torch.clamp(torch.round(w / self.scale) + self.zero, 0, self.maxq)

    def quantize(self, w: torch.Tensor) -> torch.Tensor:
        return self.scale * (self.quantize_int(w) - self.zero)


class GPTQ:
    """Collect the layer-input Hessian with `add_batch`, then `fasterquant()` rewrites `layer.weight` with the
    quantised-dequantised weights and returns (scales, zeros, g_idx) for packing."""

    def __init__(self, layer: nn.Linear) -> None:
        self.layer = layer
        self.rows, self.columns = layer.weight.shape
        self.H = torch.zeros(self.columns, self.columns, dtype=torch.float32, device=layer.weight.device)
        self.nsamples = 0

    @torch.no_grad()
    def add_batch(self, inp: torch.Tensor) -> None:
        x = inp.reshape(-1, inp.shape[-1]).t().float()            # [columns, n]
        n = x.shape[1]
        self.H *= self.nsamples / (self.nsamples + n)
        self.nsamples += n
        x = math.sqrt(2.0 / self.nsamples) * x
        self.H += x @ x.t()

    @torch.no_grad()
    def fasterquant(self, bits: int = 4, group_size: int = 128, blocksize: int = 128, percdamp: float = 0.01,
                    sym: bool = False, actorder: bool = False):
        W = self.layer.weight.data.clone().float()
        H = self.H.clone()
        quantizer = Quantizer(bits, sym)
        dead = torch.diag(H) == 0
        H[dead, dead] = 1.0
        W[:, dead] = 0.0
        perm = None
        if actorder:                                             # quantise the most "important" columns first
            perm = torch.argsort(torch.diag(H), descending=True)
            W, H = W[:, perm], H[perm][:, perm]
        damp = percdamp * torch.mean(torch.diag(H))
        H += torch.eye(self.columns, device=H.device) * damp
        Hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(H)), upper=True)
        gs = group_size if group_size > 0 else self.columns
        n_groups = math.ceil(self.columns / gs)
        scales = torch.zeros(self.rows, n_groups, device=W.device)
        zeros = torch.zeros(self.rows, n_groups, device=W.device)
        Q = torch.zeros_like(W)
        for i1 in range(0, self.columns, blocksize):
            i2 = min(i1 + blocksize, self.columns)
            W1 = W[:, i1:i2].clone()
            Err1 = torch.zeros_like(W1)
            Hinv1 = Hinv[i1:i2, i1:i2]
            for i in range(i2 - i1):
                col = i1 + i
                if col % gs == 0:
                    quantizer.find_params(W[:, col:col + gs])
                    scales[:, col // gs] = quantizer.scale[:, 0]
                    zeros[:, col // gs] = quantizer.zero[:, 0]
                w = W1[:, i]
                q = quantizer.quantize(w.unsqueeze(1)).squeeze(1)
                Q[:, col] = q
                err = (w - q) / Hinv1[i, i]
                W1[:, i:] -= err.unsqueeze(1) * Hinv1[i, i:].unsqueeze(0)
                W[:, col:i2] = W1[:, i:]                         # keep W current for the next group's min/max
                Err1[:, i] = err
            W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:]
        g_idx = torch.arange(self.columns, device=W.device) // gs
        if perm is not None:
            inv = torch.argsort(perm)
            Q = Q[:, inv]
            g_idx = g_idx[inv]
        self.layer.weight.data = Q.to(self.layer.weight.dtype)
        return scales, zeros, g_idx.to(torch.int32)


def pack_rows(q: torch.Tensor, bits: int) -> torch.Tensor:
    """Pack unsigned ints `[K, N]` along K into int32 words `[K * bits / 32, N]` (GPTQ `qweight` layout)."""
    per = 32 // bits
    K, N = q.shape
    assert K % per == 0
    q = q.to(torch.int64).reshape(K // per, per, N)
    shifts = (torch.arange(per, device=q.device) * bits).view(1, per, 1)
    word = (q << shifts).sum(1)
    word = torch.where(word >= 2 ** 31, word - 2 ** 32, word)
    return word.to(torch.int32)


def unpack_rows(packed: torch.Tensor, bits: int) -> torch.Tensor:
    per = 32 // bits
    w = packed.to(torch.int64) & 0xFFFFFFFF
    shifts = (torch.arange(per, device=packed.device) * bits).view(1, per, 1)
    out = (w.unsqueeze(1) >> shifts) & (2 ** bits - 1)
    return out.reshape(packed.shape[0] * per, packed.shape[1])


class QuantLinear(nn.Module):
    """Weight-only quantised linear: `qweight` int32 [in*bits/32, out], `qzeros` int32 [groups, out*bits/32],
    `scales` [groups, out], `g_idx` int32 [in].  Forward de-quantises into the activation dtype and runs the GEMM
    (the de-quantised tile stays in L2 / registers on the CUDA path of `ops.gemm`)."""

    def __init__(self, bits: int, group_size: int, in_features: int, out_features: int, bias: bool = True) -> None:
        super().__init__()
        assert bits in (2, 4, 8)
        self.bits, self.in_features, self.out_features = bits, in_features, out_features
        self.group_size = group_size if group_size > 0 else in_features
        groups = math.ceil(in_features / self.group_size)
        self.register_buffer("qweight", torch.zeros(in_features * bits // 32, out_features, dtype=torch.int32))
        self.register_buffer("qzeros", torch.zeros(groups, out_features * bits // 32, dtype=torch.int32))
        self.register_buffer("scales", torch.zeros(groups, out_features, dtype=torch.float16))
        self.register_buffer("g_idx", torch.arange(in_features, dtype=torch.int32) // self.group_size)
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=torch.float16)) if bias else None
        self._cache = None

    @torch.no_grad()
    def pack(self, linear: nn.Linear, scales: torch.Tensor, zeros: torch.Tensor, g_idx: Optional[torch.Tensor] = None):
        """`linear.weight` holds the quantised-dequantised weights produced by the solver."""
        if g_idx is not None:
            self.g_idx = g_idx.to(torch.int32).to(self.g_idx.device)
        W = linear.weight.data.float()                                        # [out, in]
        s = scales.float()[:, self.g_idx.long()]                              # [out, in]
        z = zeros.float()[:, self.g_idx.long()]
        q = torch.clamp(torch.round(W / s + z), 0, 2 ** self.bits - 1).t().contiguous()   # [in, out]
        self.qweight = pack_rows(q, self.bits)
        # zeros [out, groups] packed along the OUT dim -> qzeros [groups, out * bits / 32]
        self.qzeros = pack_rows(zeros.round().clamp(0, 2 ** self.bits - 1).contiguous(), self.bits).t().contiguous()
        self.scales = scales.t().contiguous().to(self.scales.dtype)
        if linear.bias is not None and self.bias is not None:
            self.bias.data = linear.bias.data.to(self.bias.dtype)
        self._cache = None

    def dequantize(self, dtype=torch.float16) -> torch.Tensor:
        q = unpack_rows(self.qweight, self.bits).to(dtype)                    # [in, out]
        z = unpack_rows(self.qzeros.t().contiguous(), self.bits).t().to(dtype)  # [groups, out]
        g = self.g_idx.long()
        return ((q - z[g]) * self.scales.to(dtype)[g]).t().contiguous()       # [out, in]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._cache is None or self._cache.dtype != x.dtype or self._cache.device != x.device:
            w = self.dequantize(x.dtype if x.dtype.is_floating_point else torch.float16)
            if not self.training:
                self._cache = w
        else:
            w = self._cache
        return nn.functional.linear(x, w, None if self.bias is None else self.bias.to(x.dtype))

    def extra_repr(self) -> str:
        return f"in={self.in_features}, out={self.out_features}, bits={self.bits}, group_size={self.group_size}"


CaiQuantLinear = QuantLinear


@torch.no_grad()
def gptq_quantize_model(model: nn.Module, calibration_batches: Iterable[Dict[str, torch.Tensor]], bits: int = 4,
                        group_size: int = 128, percdamp: float = 0.01, actorder: bool = False,
                        skip: Iterable[str] = ("lm_head",)) -> nn.Module:
    """Quantise every `nn.Linear` (except `skip`) with GPTQ using inputs observed on `calibration_batches`
    (each a dict of model kwargs) and replace it by a packed `QuantLinear`."""
    targets = {n: m for n, m in model.named_modules()
               if isinstance(m, nn.Linear) and not any(s in n for s in skip) and m.in_features % (32 // bits) == 0}
    solvers = {n: GPTQ(m) for n, m in targets.items()}
    hooks = [m.register_forward_hook(lambda mod, inp, out, n=n: solvers[n].add_batch(inp[0].detach()))
             for n, m in targets.items()]
    was_training = model.training
    model.eval()
    for batch in calibration_batches:
        model(**batch)
    for h in hooks:
        h.remove()
    for name, lin in targets.items():
        gs = group_size if (group_size > 0 and lin.in_features % group_size == 0) else -1
        scales, zeros, g_idx = solvers[name].fasterquant(bits=bits, group_size=gs, percdamp=percdamp,
                                                         actorder=actorder)
        ql = QuantLinear(bits, gs, lin.in_features, lin.out_features, bias=lin.bias is not None)
        ql = ql.to(lin.weight.device)
        ql.pack(lin, scales, zeros, g_idx)
        parent = model
        *path, leaf = name.split(".")
        for p in path:
            parent = getattr(parent, p)
        setattr(parent, leaf, ql)
    model.train(was_training)
    return model
