"""DistGaloreAwamW: GaLore AdamW for TP / ZeRO-sharded parameters — the gradient shard is gathered to its global 2-D
shape for the (periodic) SVD projection and the projected update is re-sharded.
Parity: reference `colossalai/nn/optimizer/distributed_galore.py`."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from ...interface.optimizer import DistributedOptim
from ...parallel import comm
from .galore import GaLoreProjector, _dequant, _quant

__all__ = ["DistGaloreAwamW"]


class DistGaloreAwamW(DistributedOptim):
    def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, nbits: int = 8,
                 min_8bit_size: int = 4096, percentile_clipping: int = 100, block_wise: bool = True,
                 is_paged: bool = False) -> None:
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.nbits, self.min_8bit_size = nbits, min_8bit_size
        self._post_cast()

    def _post_cast(self) -> None:
        for a, v in (("tp_size", 1), ("tp_group", None), ("dp_size", 1), ("dp_group", None), ("is_zero", False),
                     ("shard_to_working_param", {}), ("nbits", 8), ("min_8bit_size", 4096)):
            if not hasattr(self, a):
                setattr(self, a, v)

    def setup_distributed(self, tp_group=None, dp_group=None, shard_to_working_param: Optional[Dict] = {},
                          padding_map=None, is_zero: Optional[bool] = False) -> None:
        self.tp_group, self.dp_group = tp_group, dp_group
        self.tp_size = dist.get_world_size(tp_group) if tp_group is not None else 1
        self.dp_size = dist.get_world_size(dp_group) if dp_group is not None else 1
        self.shard_to_working_param = shard_to_working_param or {}
        self.is_zero = bool(is_zero)

    def _shard(self, p):
        wp = self.shard_to_working_param.get(id(p), p)
        sh = getattr(wp, "dist_shard", None)
        if sh is None and hasattr(wp, "tp_shard_dim"):       # fused (customised) column / row parallel weights
            sh = (wp.tp_shard_dim, None)
        return sh if (sh is not None and self.tp_size > 1) else None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                st.setdefault("step", 0)
                grad = p.grad.float()
                sh = self._shard(p)
                projected = "rank" in group and grad.dim() == 2
                if projected:
                    wp = self.shard_to_working_param.get(id(p), p)
                    if sh is None:
                        full = grad
                    elif hasattr(wp, "gather_fn"):               # fused q|k|v / gate|up shards: block-aware gather
                        full = wp.gather_fn(grad)
                    else:
                        full = comm.all_gather(grad, sh[0], sh[1] if sh[1] is not None else self.tp_group)
                    if "projector" not in st:
                        st["projector"] = GaLoreProjector(group["rank"], group.get("update_proj_gap", 200),
                                                          group.get("scale", 0.25), group.get("proj_type", "std"))
                    grad = st["projector"].project(full, st["step"])
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
                if projected:
                    upd = st["projector"].project_back(upd)
                    if sh is not None:
                        wp = self.shard_to_working_param.get(id(p), p)
                        upd = wp.shard_fn(upd) if hasattr(wp, "shard_fn") else \
                            comm.split_along(upd, sh[0], sh[1] if sh[1] is not None else self.tp_group)
                pf = p.float()
                if group["weight_decay"] > 0:
                    pf = pf * (1 - group["lr"] * group["weight_decay"])
                p.copy_(pf - group["lr"] * upd)
        return loss
