"""Distrifusion-style PATCH parallelism for diffusion transformers: every rank denoises one horizontal slab of the
latent.  Self-attention needs the keys / values of all slabs; because consecutive denoising steps are highly similar
the remote K/V (and conv halos) of the PREVIOUS step are good enough after a few warm-up steps, so the exchange of the
fresh tensors is launched asynchronously and only consumed one step later — communication leaves the critical path.

Parity: reference `colossalai/inference/modeling/layers/distrifusion.py:1-626` (`DistrifusionPatchEmbed`,
`DistrifusionConv2D`, `DistriSelfAttention`, `DistrifusionFusedAttention`, the patched PixArt / SD3 forwards) driven by
`PatchParallelismConfig`-like settings (`inference/config.py`: patched_parallelism_size, warm-up steps, sync mode).

Layout choices for NVSwitch: the K/V exchange is ONE flat all-gather per layer per step (K and V packed in a single
buffer), issued with `async_op=True` on the NCCL stream and waited for at the same layer of the next step.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ....parallel import comm

__all__ = ["PatchParallelContext", "DistriKVExchange", "DistriConv2d", "enable_patch_parallel",
           "disable_patch_parallel"]


@dataclass
class PatchParallelContext:
    """Shared state of one patch-parallel model: group, step counter, sync/stale mode."""

    group: Optional[dist.ProcessGroup] = None
    warmup_steps: int = 4            # steps that exchange synchronously (early steps change the most)
    mode: str = "stale"              # "stale" (async, previous-step remote K/V) | "sync" (always fresh; exact)
    step: int = 0
    exchanges: List["DistriKVExchange"] = field(default_factory=list)

    @property
    def size(self) -> int:
        return comm.group_size(self.group)

    @property
    def rank(self) -> int:
        return comm.group_rank(self.group)

    @property
    def synchronous(self) -> bool:
        return self.mode == "sync" or self.step < self.warmup_steps

    def split_rows(self, latent: torch.Tensor, patch: int) -> Tuple[torch.Tensor, int]:
        """This rank's slab of latent rows (a multiple of the patch size) and its offset in PATCH rows."""
        H = latent.shape[-2]
        assert H % (patch * self.size) == 0, f"latent height {H} not divisible by patch {patch} x ranks {self.size}"
        rows = H // self.size
        return latent[..., self.rank * rows:(self.rank + 1) * rows, :].contiguous(), self.rank * rows // patch

    def gather_rows(self, out: torch.Tensor) -> torch.Tensor:
        if self.size == 1:
            return out
        parts = [torch.empty_like(out) for _ in range(self.size)]
        dist.all_gather(parts, out.contiguous(), group=self.group)
        return torch.cat(parts, dim=-2)

    def next_step(self) -> None:
        self.step += 1

    def reset(self) -> None:
        self.step = 0
        for e in self.exchanges:
            e.reset()


class DistriKVExchange:
    """Callable plugged into a self-attention module: `(k_local, v_local, token_dim) -> (k_all, v_all)`.

    sync mode: all-gather now.  stale mode: return [stale remote slabs with the FRESH local slab spliced in], then
    launch the async all-gather of the fresh local K/V whose result becomes the stale buffer of the next step."""

    def __init__(self, ctx: PatchParallelContext) -> None:
        self.ctx = ctx
        self._buf: Optional[torch.Tensor] = None      # [size, 2, *kv_shape] gathered last step
        self._pending = None
        ctx.exchanges.append(self)

    def reset(self) -> None:
        self._wait()
        self._buf = None

    def _wait(self) -> None:
        if self._pending is not None:
            self._pending.wait()
            self._pending = None

    def __call__(self, k: torch.Tensor, v: torch.Tensor, token_dim: int = 1):
        ctx = self.ctx
        n = ctx.size
        if n == 1:
            return k, v
        local = torch.stack([k, v], 0).contiguous()
        if ctx.synchronous or self._buf is None:
            self._wait()
            buf = torch.empty((n,) + tuple(local.shape), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(buf.view(-1), local.view(-1), group=ctx.group)
            self._buf = buf
            full = buf
        else:
            self._wait()                               # the gather launched at this layer one step ago
            full = self._buf.clone()
            full[ctx.rank] = local                     # own slab is always fresh
            nxt = torch.empty_like(self._buf)
            self._pending = dist.all_gather_into_tensor(nxt.view(-1), local.view(-1), group=ctx.group, async_op=True)
            self._buf = nxt
        ks = torch.cat([full[r, 0] for r in range(n)], dim=token_dim)
        vs = torch.cat([full[r, 1] for r in range(n)], dim=token_dim)
        return ks, vs


class DistriConv2d(nn.Module):
    """Row-sharded convolution with halo exchange (UNet / VAE style convs under patch parallelism): each rank holds a
    slab of rows and needs `padding` boundary rows from its neighbours — fresh in sync mode, one step stale otherwise."""

    def __init__(self, conv: nn.Conv2d, ctx: PatchParallelContext) -> None:
        super().__init__()
        assert conv.stride[0] == 1 or conv.padding[0] == 0, "halo exchange is implemented for stride-1 convs"
        self.conv, self.ctx = conv, ctx
        self.halo = conv.padding[0]
        self._stale: Optional[torch.Tensor] = None
        self._pending = None

    def _exchange(self, x: torch.Tensor, async_op: bool):
        ctx, h = self.ctx, self.halo
        edges = torch.cat([x[..., :h, :], x[..., -h:, :]], dim=-2).contiguous()      # my top and bottom rows
        buf = torch.empty((ctx.size,) + tuple(edges.shape), dtype=x.dtype, device=x.device)
        work = dist.all_gather_into_tensor(buf.view(-1), edges.view(-1), group=ctx.group, async_op=async_op)
        return buf, work

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ctx, h = self.ctx, self.halo
        if ctx.size == 1 or h == 0:
            return self.conv(x)
        if self._pending is not None:               # the exchange launched at this layer one step ago
            self._pending.wait()
            self._pending = None
        if ctx.synchronous or self._stale is None:
            buf, _ = self._exchange(x, False)
            self._stale = buf
        else:
            buf = self._stale
            self._stale, self._pending = self._exchange(x, True)
        r, n = ctx.rank, ctx.size
        top = buf[r - 1][..., -h:, :] if r > 0 else x.new_zeros(x.shape[:-2] + (h, x.shape[-1]))
        bot = buf[r + 1][..., :h, :] if r < n - 1 else x.new_zeros(x.shape[:-2] + (h, x.shape[-1]))
        xp = torch.cat([top, x, bot], dim=-2)
        pw = self.conv.padding[1]
        xp = F.pad(xp, (pw, pw, 0, 0))
        return F.conv2d(xp, self.conv.weight, self.conv.bias, self.conv.stride, 0, self.conv.dilation, self.conv.groups)


def enable_patch_parallel(model: nn.Module, group: Optional[dist.ProcessGroup] = None, warmup_steps: int = 4,
                          mode: str = "stale") -> PatchParallelContext:
    """Turn a DiT (`models.dit.PixArtTransformer2D` / `SD3Transformer2D`) into its patch-parallel form: the model
    splits / gathers latent rows itself (`model.patch_parallel`), every image self-attention gets a K/V exchange."""
    from ....models.dit import JointBlock, PixArtBlock

    ctx = PatchParallelContext(group=group, warmup_steps=warmup_steps, mode=mode)
    model.patch_parallel = ctx
    for m in model.modules():
        if isinstance(m, PixArtBlock):
            m.attn.kv_exchange = DistriKVExchange(ctx)
        elif isinstance(m, JointBlock):
            m.kv_exchange = DistriKVExchange(ctx)
    return ctx


def disable_patch_parallel(model: nn.Module) -> None:
    from ....models.dit import JointBlock, PixArtBlock

    ctx = getattr(model, "patch_parallel", None)
    if ctx is not None:
        ctx.reset()
    model.patch_parallel = None
    for m in model.modules():
        if isinstance(m, PixArtBlock):
            m.attn.kv_exchange = None
        elif isinstance(m, JointBlock):
            m.kv_exchange = None
