"""GaLore AdamW (gradient low-rank projection) with optional 8-bit block-wise quantised moments.

Parity: reference `colossalai/nn/optimizer/galore.py` (`GaLoreAdamW8bit`, `get_galore_param_groups`).  The reference
relies on bitsandbytes for the 8-bit state; here the moments are quantised block-wise (absmax int8, block 256) in plain
PyTorch so the optimizer has no external dependency.
"""
from __future__ import annotations

import warnings
from typing import List, Optional

import torch
import torch.nn as nn
from torch.optim import Optimizer

__all__ = ["GaLoreAdamW8bit", "GaLoreProjector", "get_galore_param_groups"]


def get_galore_param_groups(model: nn.Module, weight_decay: float, rank: int = 256, update_proj_gap: int = 200,
                            scale: float = 0.25, proj_type: str = "std") -> List[dict]:
    """Put attention / MLP 2-D weights into a low-rank projected group; everything else in a plain group."""
    galore, plain = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if p.dim() == 2 and any(k in name for k in ("attn", "mlp", "proj", "fc", "dense")) and min(p.shape) > rank:
            galore.append(p)
        else:
            plain.append(p)
    return [{"params": plain, "weight_decay": weight_decay},
            {"params": galore, "rank": rank, "update_proj_gap": update_proj_gap, "scale": scale,
             "proj_type": proj_type, "weight_decay": weight_decay}]


class GaLoreProjector:
    def __init__(self, rank: int, update_proj_gap: int = 200, scale: float = 1.0, proj_type: str = "std") -> None:
        self.rank, self.update_proj_gap, self.scale, self.proj_type = rank, update_proj_gap, scale, proj_type
        self.ortho = None
        self.side = None

    def _svd(self, g: torch.Tensor, side: str) -> torch.Tensor:
        u, _, vh = torch.linalg.svd(g.float(), full_matrices=False)
        return u[:, : self.rank] if side == "left" else vh[: self.rank, :]

    def project(self, g: torch.Tensor, it: int) -> torch.Tensor:
        side = "right" if (self.proj_type == "std" and g.shape[0] >= g.shape[1]) or self.proj_type == "right" else "left"
        if self.ortho is None or it % self.update_proj_gap == 0:
            self.ortho = self._svd(g, side)
            self.side = side
        return g.float() @ self.ortho.t() if self.side == "right" else self.ortho.t() @ g.float()

    def project_back(self, low: torch.Tensor) -> torch.Tensor:
        full = low @ self.ortho if self.side == "right" else self.ortho @ low
        return full * self.scale


def _quant(x: torch.Tensor, block: int = 256):
    flat = x.reshape(-1)
    pad = (-flat.numel()) % block
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad)])
    b = flat.view(-1, block)
    absmax = b.abs().amax(dim=1, keepdim=True).clamp(min=1e-12)
    return (b / absmax * 127).round().to(torch.int8), absmax, x.shape, pad


def _dequant(q, absmax, shape, pad):
    flat = (q.float() / 127 * absmax).reshape(-1)
    if pad:
        flat = flat[:-pad]
    return flat.view(shape)


class GaLoreAdamW8bit(Optimizer):
    def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, nbits: int = 8,
                 min_8bit_size: int = 4096, percentile_clipping: int = 100, block_wise: bool = True,
                 is_paged: bool = False) -> None:
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.nbits, self.min_8bit_size = nbits, min_8bit_size

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "step" not in st:
                    st["step"] = 0
                grad = p.grad.float()
                if "rank" in group and grad.dim() == 2:
                    if "projector" not in st:
                        st["projector"] = GaLoreProjector(group["rank"], group.get("update_proj_gap", 200),
                                                          group.get("scale", 0.25), group.get("proj_type", "std"))
                    grad = st["projector"].project(grad, st["step"])
                st["step"] += 1
                use8 = self.nbits == 8 and grad.numel() >= self.min_8bit_size
                if "m" not in st:
                    z = torch.zeros_like(grad)
                    st["m"], st["v"] = (_quant(z), _quant(z)) if use8 else (z, z.clone())
                m = _dequant(*st["m"]) if use8 else st["m"]
                v = _dequant(*st["v"]) if use8 else st["v"]
                m.mul_(b1).add_(grad, alpha=1 - b1)
                v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
                bc1, bc2 = 1 - b1 ** st["step"], 1 - b2 ** st["step"]
                upd = (m / bc1) / ((v / bc2).sqrt() + group["eps"])
                if use8:
                    st["m"], st["v"] = _quant(m), _quant(v)
                if "projector" in st:
                    upd = st["projector"].project_back(upd)
                pf = p.float()
                if group["weight_decay"] > 0:
                    pf = pf * (1 - group["lr"] * group["weight_decay"])
                p.copy_(pf - group["lr"] * upd)
        return loss
