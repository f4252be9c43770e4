"""Segment Anything: windowed-ViT image encoder with decomposed relative positions + conv neck, prompt encoder
(random-Fourier positional encoding, point / box / mask prompts), two-way-transformer mask decoder with
hyper-network mask heads and an IoU head.

Parity: reference `colossalai/shardformer/policies/sam.py:14-260` + `modeling/sam.py:8-220` (`SamModel`).  The
vision attention's q/k/v/out projections, the MLPs and the mask-decoder attention projections are the
tensor-parallel surfaces.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..parallel import comm
from ..shardformer.layer._operation import reduce_backward
from .encdec import init_module_weights

__all__ = ["SamConfig", "SamVisionAttention", "SamVisionLayer", "SamVisionEncoder", "SamPromptEncoder",
           "SamAttention", "SamTwoWayBlock", "SamMaskDecoder", "SamModel", "SAM_ZOO"]


@dataclass
class SamConfig:
    model_type: str = "sam"
    image_size: int = 1024
    patch_size: int = 16
    vision_hidden_size: int = 768
    vision_layers: int = 12
    vision_heads: int = 12
    vision_mlp_dim: int = 3072
    window_size: int = 14
    global_attn_indexes: Tuple[int, ...] = (2, 5, 8, 11)
    use_rel_pos: bool = True
    output_channels: int = 256
    layer_norm_eps: float = 1e-6
    # prompt encoder / mask decoder
    num_point_embeddings: int = 4
    mask_input_channels: int = 16
    decoder_layers: int = 2
    decoder_heads: int = 8
    decoder_mlp_dim: int = 2048
    attention_downsample_rate: int = 2
    num_multimask_outputs: int = 3
    iou_head_depth: int = 3
    iou_head_hidden_dim: int = 256
    initializer_range: float = 0.02

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def hidden_size(self) -> int:
        return self.vision_hidden_size

    def replace(self, **kw) -> "SamConfig":
        return replace(self, **kw)


SAM_ZOO: Dict[str, SamConfig] = {
    "sam-vit-base": SamConfig(),
    "sam-vit-large": SamConfig(vision_hidden_size=1024, vision_layers=24, vision_heads=16, vision_mlp_dim=4096,
                               global_attn_indexes=(5, 11, 17, 23)),
    "sam-vit-huge": SamConfig(vision_hidden_size=1280, vision_layers=32, vision_heads=16, vision_mlp_dim=5120,
                              global_attn_indexes=(7, 15, 23, 31)),
    "sam-tiny": SamConfig(image_size=64, patch_size=8, vision_hidden_size=32, vision_layers=2, vision_heads=2,
                          vision_mlp_dim=64, window_size=4, global_attn_indexes=(1,), output_channels=16,
                          mask_input_channels=4, decoder_heads=2, decoder_mlp_dim=32, iou_head_hidden_dim=16),
}


class LayerNorm2d(nn.Module):
    """LayerNorm over the channel dim of an NCHW map."""

    def __init__(self, channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        return (x - u) / torch.sqrt(s + self.eps) * self.weight[:, None, None] + self.bias[:, None, None]


def _rel_table(q_size: int, k_size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """Pick the `[q, k, D]` slice of a `[2*max-1, D]` relative-position table (interpolated if sizes differ)."""
    need = 2 * max(q_size, k_size) - 1
    if rel_pos.shape[0] != need:
        rel_pos = F.interpolate(rel_pos.t()[None], size=need, mode="linear")[0].t()
    q = torch.arange(q_size, device=rel_pos.device)[:, None] * max(k_size / q_size, 1.0)
    k = torch.arange(k_size, device=rel_pos.device)[None, :] * max(q_size / k_size, 1.0)
    idx = (q - k) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rel_pos[idx.long()]


class SamVisionAttention(nn.Module):
    def __init__(self, cfg: SamConfig, input_size: int) -> None:
        super().__init__()
        H = cfg.vision_hidden_size
        self.num_heads, self.head_dim = cfg.vision_heads, H // cfg.vision_heads
        self.qkv_proj = nn.Linear(H, 3 * H)
        self.o_proj = nn.Linear(H, H)
        self.use_rel_pos = cfg.use_rel_pos
        if cfg.use_rel_pos:
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size - 1, self.head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size - 1, self.head_dim))
        self.shard_config = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """`x` [B, Hh, Ww, C] (a window or the full grid)."""
        B, Hh, Ww, _ = x.shape
        D = self.head_dim
        qkv = self.qkv_proj(x).reshape(B, Hh * Ww, -1)
        h = qkv.shape[-1] // (3 * D)
        q, k, v = (t.reshape(B, Hh * Ww, h, D).transpose(1, 2) for t in qkv.split(h * D, dim=-1))
        bias = None
        if self.use_rel_pos:
            rh, rw = self.rel_pos_h, self.rel_pos_w
            sc = self.shard_config
            if sc is not None and sc.enable_tensor_parallelism and comm.group_size(sc.tensor_parallel_process_group) > 1:
                # the tables are replicated but each rank only sees its heads: sum the partial grads over TP
                rh = reduce_backward(rh, sc.tensor_parallel_process_group)
                rw = reduce_backward(rw, sc.tensor_parallel_process_group)
            Rh = _rel_table(Hh, Hh, rh).to(q.dtype)
            Rw = _rel_table(Ww, Ww, rw).to(q.dtype)
            r_q = q.reshape(B, h, Hh, Ww, D)
            rel_h = torch.einsum("bnhwc,hkc->bnhwk", r_q, Rh)
            rel_w = torch.einsum("bnhwc,wkc->bnhwk", r_q, Rw)
            bias = (rel_h[..., :, None] + rel_w[..., None, :]).reshape(B, h, Hh * Ww, Hh * Ww)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias)
        return self.o_proj(o.transpose(1, 2).reshape(B, Hh, Ww, h * D))


class SamMLP(nn.Module):
    def __init__(self, hidden: int, mlp_dim: int, act: str = "gelu") -> None:
        super().__init__()
        self.up_proj = nn.Linear(hidden, mlp_dim)
        self.down_proj = nn.Linear(mlp_dim, hidden)
        self.act = F.gelu if act == "gelu" else F.relu

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.down_proj(self.act(self.up_proj(x)))


def _window_partition(x: torch.Tensor, w: int):
    B, H, W, C = x.shape
    ph, pw = (w - H % w) % w, (w - W % w) % w
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.reshape(B, Hp // w, w, Wp // w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, w, w, C)
    return x, (Hp, Wp)


def _window_unpartition(win: torch.Tensor, w: int, pad_hw, hw):
    Hp, Wp = pad_hw
    H, W = hw
    B = win.shape[0] // (Hp * Wp // w // w)
    x = win.reshape(B, Hp // w, Wp // w, w, w, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W]


class SamVisionLayer(nn.Module):
    def __init__(self, cfg: SamConfig, window_size: int) -> None:
        super().__init__()
        H = cfg.vision_hidden_size
        self.window_size = window_size
        self.norm1 = nn.LayerNorm(H, eps=cfg.layer_norm_eps)
        self.attn = SamVisionAttention(cfg, window_size if window_size > 0 else cfg.grid)
        self.norm2 = nn.LayerNorm(H, eps=cfg.layer_norm_eps)
        self.mlp = SamMLP(H, cfg.vision_mlp_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        res = x
        x = self.norm1(x)
        if self.window_size > 0:
            hw = x.shape[1:3]
            x, pad_hw = _window_partition(x, self.window_size)
            x = _window_unpartition(self.attn(x), self.window_size, pad_hw, hw)
        else:
            x = self.attn(x)
        x = res + x
        return x + self.mlp(self.norm2(x))


class SamVisionEncoder(nn.Module):
    def __init__(self, cfg: SamConfig) -> None:
        super().__init__()
        self.cfg = cfg
        H = cfg.vision_hidden_size
        self.patch_embed = nn.Conv2d(3, H, cfg.patch_size, stride=cfg.patch_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, cfg.grid, cfg.grid, H))
        self.layers = nn.ModuleList([
            SamVisionLayer(cfg, 0 if i in cfg.global_attn_indexes else cfg.window_size)
            for i in range(cfg.vision_layers)])
        C = cfg.output_channels
        self.neck = nn.Sequential(nn.Conv2d(H, C, 1, bias=False), LayerNorm2d(C),
                                  nn.Conv2d(C, C, 3, padding=1, bias=False), LayerNorm2d(C))
        self.gradient_checkpointing = False

    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        x = self.patch_embed(pixel_values.to(self.patch_embed.weight.dtype)).permute(0, 2, 3, 1)
        x = x + self.pos_embed.to(x.dtype)
        for blk in self.layers:
            if self.gradient_checkpointing and self.training:
                x = torch.utils.checkpoint.checkpoint(blk, x, use_reentrant=False)
            else:
                x = blk(x)
        return self.neck(x.permute(0, 3, 1, 2))        # [B, C, grid, grid]


class SamPositionalEmbedding(nn.Module):
    def __init__(self, channels: int, scale: float = 1.0) -> None:
        super().__init__()
        self.register_buffer("positional_embedding", scale * torch.randn(2, channels // 2))

    def encode(self, coords: torch.Tensor) -> torch.Tensor:
        """`coords` in [0, 1], last dim 2."""
        c = (2 * coords - 1) @ self.positional_embedding.to(coords.dtype)
        c = 2 * math.pi * c
        return torch.cat([c.sin(), c.cos()], dim=-1)

    def grid(self, size: int, device, dtype) -> torch.Tensor:
        ones = torch.ones(size, size, device=device, dtype=dtype)
        y = (ones.cumsum(0) - 0.5) / size
        x = (ones.cumsum(1) - 0.5) / size
        return self.encode(torch.stack([x, y], dim=-1)).permute(2, 0, 1)[None]       # [1, C, size, size]


class SamPromptEncoder(nn.Module):
    def __init__(self, cfg: SamConfig) -> None:
        super().__init__()
        self.cfg = cfg
        C = cfg.output_channels
        self.shared_embedding = SamPositionalEmbedding(C)
        self.point_embed = nn.Embedding(cfg.num_point_embeddings, C)    # neg point, pos point, box corner 1, corner 2
        self.not_a_point_embed = nn.Embedding(1, C)
        self.no_mask_embed = nn.Embedding(1, C)
        m = cfg.mask_input_channels
        self.mask_embed = nn.Sequential(
            nn.Conv2d(1, m // 4, 2, stride=2), LayerNorm2d(m // 4), nn.GELU(),
            nn.Conv2d(m // 4, m, 2, stride=2), LayerNorm2d(m), nn.GELU(), nn.Conv2d(m, C, 1))

    def forward(self, points: Optional[torch.Tensor], labels: Optional[torch.Tensor], boxes: Optional[torch.Tensor],
                masks: Optional[torch.Tensor], batch: int, device, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
        cfg = self.cfg
        C, S = cfg.output_channels, float(cfg.image_size)
        sparse = torch.zeros(batch, 0, C, device=device, dtype=dtype)
        if points is not None:
            pts = points.to(dtype) + 0.5
            lab = labels if labels is not None else torch.ones(points.shape[:2], dtype=torch.long, device=device)
            if boxes is None:          # pad with a "not a point" so the token count is prompt-type independent
                pts = torch.cat([pts, torch.zeros(batch, 1, 2, device=device, dtype=dtype)], dim=1)
                lab = torch.cat([lab, -torch.ones(batch, 1, dtype=lab.dtype, device=device)], dim=1)
            emb = self.shared_embedding.encode(pts / S)
            is_pad = (lab == -1)[..., None]
            emb = torch.where(is_pad, self.not_a_point_embed.weight[0].to(dtype).expand_as(emb), emb)
            kind = self.point_embed(lab.clamp(min=0)).to(dtype)
            emb = emb + torch.where(is_pad, torch.zeros_like(kind), kind)
            sparse = torch.cat([sparse, emb], dim=1)
        if boxes is not None:
            b = (boxes.to(dtype) + 0.5).reshape(batch, -1, 2, 2)
            emb = self.shared_embedding.encode(b / S)
            emb = emb + self.point_embed.weight[2:4].to(dtype)[None, None]
            sparse = torch.cat([sparse, emb.reshape(batch, -1, C)], dim=1)
        if masks is not None:
            dense = self.mask_embed(masks.to(dtype))
        else:
            dense = self.no_mask_embed.weight.to(dtype).reshape(1, C, 1, 1).expand(batch, C, cfg.grid, cfg.grid)
        return sparse, dense


class SamAttention(nn.Module):
    """Mask-decoder attention with an internal width of `hidden / downsample_rate`."""

    def __init__(self, hidden: int, heads: int, downsample_rate: int = 1) -> None:
        super().__init__()
        inner = hidden // downsample_rate
        self.num_heads, self.head_dim = heads, inner // heads
        self.q_proj = nn.Linear(hidden, inner)
        self.k_proj = nn.Linear(hidden, inner)
        self.v_proj = nn.Linear(hidden, inner)
        self.o_proj = nn.Linear(inner, hidden)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
        B, D = q.shape[0], self.head_dim
        q, k, v = self.q_proj(q), self.k_proj(k), self.v_proj(v)
        h = q.shape[-1] // D
        q, k, v = (t.reshape(B, t.shape[1], h, D).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v)
        return self.o_proj(o.transpose(1, 2).reshape(B, -1, h * D))


class SamTwoWayBlock(nn.Module):
    def __init__(self, cfg: SamConfig, skip_first_layer_pe: bool) -> None:
        super().__init__()
        C, nh, ds = cfg.output_channels, cfg.decoder_heads, cfg.attention_downsample_rate
        self.self_attn = SamAttention(C, nh, 1)
        self.norm1 = nn.LayerNorm(C)
        self.cross_attn_token_to_image = SamAttention(C, nh, ds)
        self.norm2 = nn.LayerNorm(C)
        self.mlp = SamMLP(C, cfg.decoder_mlp_dim, act="relu")
        self.norm3 = nn.LayerNorm(C)
        self.cross_attn_image_to_token = SamAttention(C, nh, ds)
        self.norm4 = nn.LayerNorm(C)
        self.skip_first_layer_pe = skip_first_layer_pe

    def forward(self, queries, keys, query_pe, key_pe):
        if self.skip_first_layer_pe:
            queries = self.self_attn(queries, queries, queries)
        else:
            q = queries + query_pe
            queries = queries + self.self_attn(q, q, queries)
        queries = self.norm1(queries)
        q, k = queries + query_pe, keys + key_pe
        queries = self.norm2(queries + self.cross_attn_token_to_image(q, k, keys))
        queries = self.norm3(queries + self.mlp(queries))
        q, k = queries + query_pe, keys + key_pe
        keys = self.norm4(keys + self.cross_attn_image_to_token(k, q, queries))
        return queries, keys


class _Head(nn.Module):
    def __init__(self, inp: int, hidden: int, out: int, depth: int, sigmoid: bool = False) -> None:
        super().__init__()
        dims = [inp] + [hidden] * (depth - 1) + [out]
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.sigmoid = sigmoid

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for i, l in enumerate(self.layers):
            x = l(x)
            if i < len(self.layers) - 1:
                x = F.relu(x)
        return torch.sigmoid(x) if self.sigmoid else x


class SamMaskDecoder(nn.Module):
    def __init__(self, cfg: SamConfig) -> None:
        super().__init__()
        self.cfg = cfg
        C = cfg.output_channels
        self.num_mask_tokens = cfg.num_multimask_outputs + 1
        self.iou_token = nn.Embedding(1, C)
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, C)
        self.layers = nn.ModuleList([SamTwoWayBlock(cfg, skip_first_layer_pe=(i == 0))
                                     for i in range(cfg.decoder_layers)])
        self.final_attn_token_to_image = SamAttention(C, cfg.decoder_heads, cfg.attention_downsample_rate)
        self.norm_final = nn.LayerNorm(C)
        self.upscale = nn.Sequential(nn.ConvTranspose2d(C, C // 4, 2, stride=2), LayerNorm2d(C // 4), nn.GELU(),
                                     nn.ConvTranspose2d(C // 4, C // 8, 2, stride=2), nn.GELU())
        self.output_hypernetworks_mlps = nn.ModuleList([_Head(C, C, C // 8, 3) for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = _Head(C, cfg.iou_head_hidden_dim, self.num_mask_tokens, cfg.iou_head_depth)

    def forward(self, image_embeddings, image_pe, sparse, dense, multimask_output: bool = True):
        B, C, H, W = image_embeddings.shape
        tokens = torch.cat([self.iou_token.weight, self.mask_tokens.weight], dim=0).to(sparse.dtype)
        tokens = torch.cat([tokens[None].expand(B, -1, -1), sparse], dim=1)
        src = (image_embeddings + dense).flatten(2).transpose(1, 2)
        pos = image_pe.expand(B, -1, -1, -1).flatten(2).transpose(1, 2)
        q, k = tokens, src
        for blk in self.layers:
            q, k = blk(q, k, tokens, pos)
        q = self.norm_final(q + self.final_attn_token_to_image(q + tokens, k + pos, k))
        iou_tok, mask_toks = q[:, 0], q[:, 1: 1 + self.num_mask_tokens]
        up = self.upscale(k.transpose(1, 2).reshape(B, C, H, W))
        hyper = torch.stack([m(mask_toks[:, i]) for i, m in enumerate(self.output_hypernetworks_mlps)], dim=1)
        b, c, h, w = up.shape
        masks = (hyper @ up.reshape(b, c, h * w)).reshape(b, -1, h, w)
        iou = self.iou_prediction_head(iou_tok)
        sl = slice(1, None) if multimask_output else slice(0, 1)
        return masks[:, sl], iou[:, sl]


class SamModel(nn.Module):
    def __init__(self, config: Optional[SamConfig] = None, **kw) -> None:
        super().__init__()
        cfg = config or SamConfig(**kw)
        self.cfg = self.config = cfg
        self.vision_encoder = SamVisionEncoder(cfg)
        self.prompt_encoder = SamPromptEncoder(cfg)
        self.mask_decoder = SamMaskDecoder(cfg)
        self.shard_config = None
        init_module_weights(self, cfg.initializer_range)

    def gradient_checkpointing_enable(self, *a, **k) -> None:
        self.vision_encoder.gradient_checkpointing = True

    def get_image_embeddings(self, pixel_values: torch.Tensor) -> torch.Tensor:
        return self.vision_encoder(pixel_values)

    def forward(self, pixel_values: Optional[torch.Tensor] = None, input_points: Optional[torch.Tensor] = None,
                input_labels: Optional[torch.Tensor] = None, input_boxes: Optional[torch.Tensor] = None,
                input_masks: Optional[torch.Tensor] = None, image_embeddings: Optional[torch.Tensor] = None,
                multimask_output: bool = True, labels: Optional[torch.Tensor] = None,
                **unused) -> Dict[str, torch.Tensor]:
        """`input_points` [B, P, 2] in pixel coordinates, `input_labels` [B, P] (1 fg / 0 bg / -1 pad),
        `input_boxes` [B, nb, 4]; `labels` (optional) are target masks [B, M, 4*grid, 4*grid] for a BCE + IoU loss."""
        emb = image_embeddings if image_embeddings is not None else self.vision_encoder(pixel_values)
        B = emb.shape[0]
        pe = self.prompt_encoder.shared_embedding.grid(self.cfg.grid, emb.device, emb.dtype)
        sparse, dense = self.prompt_encoder(input_points, input_labels, input_boxes, input_masks, B, emb.device,
                                            emb.dtype)
        masks, iou = self.mask_decoder(emb, pe, sparse, dense, multimask_output)
        out = {"pred_masks": masks, "iou_scores": iou, "image_embeddings": emb}
        if labels is not None:
            tgt = labels.to(masks.dtype)
            bce = F.binary_cross_entropy_with_logits(masks.float(), tgt.float())
            with torch.no_grad():
                pm = (masks > 0).float()
                inter = (pm * tgt).sum((-1, -2))
                union = ((pm + tgt) > 0).float().sum((-1, -2)).clamp(min=1.0)
            out["loss"] = bce + F.mse_loss(iou.float(), inter / union)
        return out
