"""Diffusion transformers for the diffusion engine: a PixArt-alpha style DiT (adaLN-single, self-attention +
text cross-attention) and an SD3-style MMDiT (joint attention over text and image streams), with the two samplers they
are used with (DDIM for epsilon prediction, Euler flow matching) and a minimal pipeline object exposing the `diffusers`
call protocol (`pipe(prompt_embeds=..., num_inference_steps=..., guidance_scale=...) -> .images`).

Parity: the reference drives `diffusers`' `PixArtAlphaPipeline` / `StableDiffusion3Pipeline`
(`colossalai/inference/modeling/models/{pixart_alpha.py:1-220, stablediffusion3.py:1-180}`,
`inference/modeling/layers/diffusion.py`); `diffusers` is not available offline, so the transformer backbones are
implemented natively here and the engine accepts either these or a real diffusers pipeline.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encdec import BlockConfig, FeedForward, MultiHeadAttention, init_module_weights

__all__ = ["DiTConfig", "PixArtTransformer2D", "SD3Transformer2D", "DDIMScheduler", "FlowMatchEulerScheduler",
           "LatentDecoder", "DiffusionPipeline", "DiffusionOutput", "DIT_ZOO", "build_diffusion_pipeline"]


@dataclass
class DiTConfig:
    model_type: str = "pixart_alpha"     # "pixart_alpha" | "stable_diffusion_3"
    sample_size: int = 64                # latent height = width
    patch_size: int = 2
    in_channels: int = 4
    hidden_size: int = 1152
    num_layers: int = 28
    num_heads: int = 16
    mlp_ratio: float = 4.0
    caption_channels: int = 4096         # text-encoder width (T5-XXL for PixArt)
    pooled_projection_dim: int = 2048    # SD3 pooled CLIP embedding
    learn_sigma: bool = True             # PixArt predicts (eps, sigma): 2 * in_channels outputs
    max_text_len: int = 120

    @property
    def out_channels(self) -> int:
        return self.in_channels * 2 if (self.learn_sigma and self.model_type == "pixart_alpha") else self.in_channels

    @property
    def grid(self) -> int:
        return self.sample_size // self.patch_size

    def block(self, cross: bool = False) -> BlockConfig:
        return BlockConfig(hidden_size=self.hidden_size, num_heads=self.num_heads,
                           ffn_dim=int(self.hidden_size * self.mlp_ratio), act="gelu_tanh", norm_eps=1e-6)

    def replace(self, **kw) -> "DiTConfig":
        return replace(self, **kw)


DIT_ZOO: Dict[str, DiTConfig] = {
    "pixart-alpha-xl-2": DiTConfig(),
    "sd3-medium": DiTConfig(model_type="stable_diffusion_3", sample_size=128, in_channels=16, hidden_size=1536,
                            num_layers=24, num_heads=24, learn_sigma=False),
    "pixart-tiny": DiTConfig(sample_size=8, patch_size=2, in_channels=4, hidden_size=32, num_layers=2, num_heads=2,
                             caption_channels=24, max_text_len=6),
    "sd3-tiny": DiTConfig(model_type="stable_diffusion_3", sample_size=8, patch_size=2, in_channels=4, hidden_size=32,
                          num_layers=2, num_heads=2, caption_channels=24, pooled_projection_dim=16, learn_sigma=False,
                          max_text_len=6),
}


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([args.cos(), args.sin()], dim=-1)


def sincos_2d(grid_h: int, grid_w: int, dim: int, row_offset: int = 0, device=None) -> torch.Tensor:
    """Fixed 2-D sin/cos position table `[grid_h * grid_w, dim]`; `row_offset` lets a patch-parallel rank build the
    rows of its own slab."""
    def one(pos, d):
        omega = 1.0 / (10000 ** (torch.arange(d // 2, dtype=torch.float32, device=device) / (d / 2)))
        out = pos.reshape(-1)[:, None] * omega[None]
        return torch.cat([out.sin(), out.cos()], dim=1)

    ys = torch.arange(row_offset, row_offset + grid_h, dtype=torch.float32, device=device)
    xs = torch.arange(grid_w, dtype=torch.float32, device=device)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.cat([one(gx, dim // 2), one(gy, dim // 2)], dim=1)


class PatchEmbed(nn.Module):
    def __init__(self, cfg: DiTConfig) -> None:
        super().__init__()
        self.cfg = cfg
        self.proj = nn.Conv2d(cfg.in_channels, cfg.hidden_size, cfg.patch_size, stride=cfg.patch_size)

    def forward(self, latent: torch.Tensor, row_offset: int = 0) -> torch.Tensor:
        x = self.proj(latent.to(self.proj.weight.dtype))
        gh, gw = x.shape[-2:]
        x = x.flatten(2).transpose(1, 2)
        return x + sincos_2d(gh, gw, x.shape[-1], row_offset, x.device).to(x.dtype)[None]


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden: int, freq_dim: int = 256) -> None:
        super().__init__()
        self.freq_dim = freq_dim
        self.linear_1 = nn.Linear(freq_dim, hidden)
        self.linear_2 = nn.Linear(hidden, hidden)

    def forward(self, t: torch.Tensor, dtype) -> torch.Tensor:
        return self.linear_2(F.silu(self.linear_1(timestep_embedding(t, self.freq_dim).to(dtype))))


def _modulate(x, shift, scale):
    return x * (1 + scale) + shift


class PixArtBlock(nn.Module):
    """adaLN-single block: the six modulation vectors come from ONE shared timestep MLP plus a per-block table."""

    def __init__(self, cfg: DiTConfig) -> None:
        super().__init__()
        bc = cfg.block()
        H = cfg.hidden_size
        self.scale_shift_table = nn.Parameter(torch.randn(6, H) / H ** 0.5)
        self.norm1 = nn.LayerNorm(H, elementwise_affine=False, eps=1e-6)
        self.attn = MultiHeadAttention(bc)
        self.cross_attn = MultiHeadAttention(bc, cross=True)
        self.norm2 = nn.LayerNorm(H, elementwise_affine=False, eps=1e-6)
        self.mlp = FeedForward(bc)

    def forward(self, x: torch.Tensor, text: torch.Tensor, t6: torch.Tensor,
                text_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = x.shape[0]
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = (self.scale_shift_table[None].to(x.dtype) + t6.reshape(B, 6, -1)).chunk(6, 1)
        x = x + g_a * self.attn(_modulate(self.norm1(x), sh_a, sc_a))
        x = x + self.cross_attn(x, memory=text, key_padding_mask=text_mask)
        return x + g_m * self.mlp(_modulate(self.norm2(x), sh_m, sc_m))


class PixArtTransformer2D(nn.Module):
    def __init__(self, config: Optional[DiTConfig] = None, **kw) -> None:
        super().__init__()
        cfg = config or DiTConfig(**kw)
        self.cfg = self.config = cfg
        H = cfg.hidden_size
        self.pos_embed = PatchEmbed(cfg)
        self.time_embed = TimestepEmbedder(H)
        self.t_block = nn.Linear(H, 6 * H)
        self.caption_projection = nn.Sequential(nn.Linear(cfg.caption_channels, H), nn.GELU(approximate="tanh"),
                                                nn.Linear(H, H))
        self.transformer_blocks = nn.ModuleList([PixArtBlock(cfg) for _ in range(cfg.num_layers)])
        self.norm_out = nn.LayerNorm(H, elementwise_affine=False, eps=1e-6)
        self.scale_shift_table = nn.Parameter(torch.randn(2, H) / H ** 0.5)
        self.proj_out = nn.Linear(H, cfg.patch_size ** 2 * cfg.out_channels)
        self.patch_parallel = None           # set by `layers.distrifusion.enable_patch_parallel`
        init_module_weights(self, 0.02)

    def unpatchify(self, x: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
        p, c = self.cfg.patch_size, self.cfg.out_channels
        x = x.reshape(x.shape[0], gh, gw, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, gh * p, gw * p)

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, timestep: torch.Tensor,
                encoder_attention_mask: Optional[torch.Tensor] = None, **unused) -> torch.Tensor:
        pp = self.patch_parallel
        lat = hidden_states
        row_offset = 0
        if pp is not None:
            lat, row_offset = pp.split_rows(lat, self.cfg.patch_size)
        x = self.pos_embed(lat, row_offset)
        gh, gw = lat.shape[-2] // self.cfg.patch_size, lat.shape[-1] // self.cfg.patch_size
        t = self.time_embed(timestep.reshape(-1).expand(x.shape[0]), x.dtype)
        t6 = self.t_block(F.silu(t))
        text = self.caption_projection(encoder_hidden_states.to(x.dtype))
        for blk in self.transformer_blocks:
            x = blk(x, text, t6, encoder_attention_mask)
        shift, scale = (self.scale_shift_table[None].to(x.dtype) + t[:, None]).chunk(2, dim=1)
        x = self.proj_out(_modulate(self.norm_out(x), shift, scale))
        out = self.unpatchify(x, gh, gw)
        if pp is not None:
            out = pp.gather_rows(out)
        return out


class JointBlock(nn.Module):
    """MMDiT block: image and text streams have their own adaLN-zero modulation, QKV and MLP but attend jointly."""

    def __init__(self, cfg: DiTConfig, last: bool = False) -> None:
        super().__init__()
        bc = cfg.block()
        H = cfg.hidden_size
        self.num_heads, self.head_dim = cfg.num_heads, H // cfg.num_heads
        self.last = last
        self.ada_x = nn.Linear(H, 6 * H)
        self.ada_c = nn.Linear(H, (2 if last else 6) * H)
        self.norm_x = nn.LayerNorm(H, elementwise_affine=False, eps=1e-6)
        self.norm_c = nn.LayerNorm(H, elementwise_affine=False, eps=1e-6)
        self.qkv_x, self.qkv_c = nn.Linear(H, 3 * H), nn.Linear(H, 3 * H)
        self.out_x = nn.Linear(H, H)
        self.norm2_x = nn.LayerNorm(H, elementwise_affine=False, eps=1e-6)
        self.mlp_x = FeedForward(bc)
        if not last:
            self.out_c = nn.Linear(H, H)
            self.norm2_c = nn.LayerNorm(H, elementwise_affine=False, eps=1e-6)
            self.mlp_c = FeedForward(bc)
        self.kv_exchange: Optional[Callable] = None     # patch parallel: image K/V of the other ranks

    def forward(self, x: torch.Tensor, c: torch.Tensor, temb: torch.Tensor):
        B, Nx, H = x.shape
        e = F.silu(temb)
        sa, ca, ga, sm, cm, gm = self.ada_x(e)[:, None].chunk(6, -1)
        mods_c = self.ada_c(e)[:, None].chunk(2 if self.last else 6, -1)
        hx = _modulate(self.norm_x(x), sa, ca)
        hc = _modulate(self.norm_c(c), mods_c[0], mods_c[1])
        qx, kx, vx = self.qkv_x(hx).chunk(3, -1)
        qc, kc, vc = self.qkv_c(hc).chunk(3, -1)
        if self.kv_exchange is not None:
            kx_all, vx_all = self.kv_exchange(kx, vx, token_dim=1)
        else:
            kx_all, vx_all = kx, vx

        def heads(t):
            return t.reshape(B, t.shape[1], self.num_heads, self.head_dim).transpose(1, 2)

        q = heads(torch.cat([qc, qx], 1))
        k = heads(torch.cat([kc, kx_all], 1))
        v = heads(torch.cat([vc, vx_all], 1))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, -1, H)
        oc, ox = o[:, : c.shape[1]], o[:, c.shape[1]:]
        x = x + ga * self.out_x(ox)
        x = x + gm * self.mlp_x(_modulate(self.norm2_x(x), sm, cm))
        if not self.last:
            c = c + mods_c[2] * self.out_c(oc)
            c = c + mods_c[5] * self.mlp_c(_modulate(self.norm2_c(c), mods_c[3], mods_c[4]))
        return x, c


class SD3Transformer2D(nn.Module):
    def __init__(self, config: Optional[DiTConfig] = None, **kw) -> None:
        super().__init__()
        cfg = config or DiTConfig(model_type="stable_diffusion_3", learn_sigma=False, **kw)
        self.cfg = self.config = cfg
        H = cfg.hidden_size
        self.pos_embed = PatchEmbed(cfg)
        self.time_embed = TimestepEmbedder(H)
        self.pooled_embed = nn.Sequential(nn.Linear(cfg.pooled_projection_dim, H), nn.SiLU(), nn.Linear(H, H))
        self.context_embedder = nn.Linear(cfg.caption_channels, H)
        self.transformer_blocks = nn.ModuleList([JointBlock(cfg, last=(i == cfg.num_layers - 1))
                                                 for i in range(cfg.num_layers)])
        self.norm_out = nn.LayerNorm(H, elementwise_affine=False, eps=1e-6)
        self.ada_out = nn.Linear(H, 2 * H)
        self.proj_out = nn.Linear(H, cfg.patch_size ** 2 * cfg.out_channels)
        self.patch_parallel = None
        init_module_weights(self, 0.02)

    unpatchify = PixArtTransformer2D.unpatchify

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, timestep: torch.Tensor,
                pooled_projections: Optional[torch.Tensor] = None, **unused) -> torch.Tensor:
        pp = self.patch_parallel
        lat, row_offset = hidden_states, 0
        if pp is not None:
            lat, row_offset = pp.split_rows(lat, self.cfg.patch_size)
        x = self.pos_embed(lat, row_offset)
        gh, gw = lat.shape[-2] // self.cfg.patch_size, lat.shape[-1] // self.cfg.patch_size
        temb = self.time_embed(timestep.reshape(-1).expand(x.shape[0]), x.dtype)
        if pooled_projections is not None:
            temb = temb + self.pooled_embed(pooled_projections.to(x.dtype))
        c = self.context_embedder(encoder_hidden_states.to(x.dtype))
        for blk in self.transformer_blocks:
            x, c = blk(x, c, temb)
        shift, scale = self.ada_out(F.silu(temb))[:, None].chunk(2, -1)
        out = self.unpatchify(self.proj_out(_modulate(self.norm_out(x), shift, scale)), gh, gw)
        if pp is not None:
            out = pp.gather_rows(out)
        return out


# ------------------------------------------------------------------------------------------------ samplers
class DDIMScheduler:
    """Deterministic DDIM over a scaled-linear beta schedule (epsilon prediction)."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02) -> None:
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, 0)
        self.num_train_timesteps = num_train_timesteps
        self.timesteps: torch.Tensor = torch.arange(num_train_timesteps - 1, -1, -1)
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n: int, device=None) -> None:
        step = self.num_train_timesteps // n
        self.timesteps = (torch.arange(0, n) * step).flip(0).to(device)
        self._step = step

    def step(self, model_output: torch.Tensor, t, sample: torch.Tensor) -> torch.Tensor:
        t = int(t)
        a_t = self.alphas_cumprod[t].to(sample.device, sample.dtype)
        prev = t - self._step
        a_p = self.alphas_cumprod[prev].to(sample.device, sample.dtype) if prev >= 0 else sample.new_tensor(1.0)
        x0 = (sample - (1 - a_t).sqrt() * model_output) / a_t.sqrt()
        return a_p.sqrt() * x0 + (1 - a_p).sqrt() * model_output


class FlowMatchEulerScheduler:
    """Rectified-flow Euler sampler (velocity prediction): x <- x + (sigma_next - sigma) * v."""

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 3.0) -> None:
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n: int, device=None) -> None:
        s = torch.linspace(1.0, 1.0 / n, n)
        s = self.shift * s / (1 + (self.shift - 1) * s)
        self.sigmas = torch.cat([s, s.new_zeros(1)]).to(device)
        self.timesteps = (s * self.num_train_timesteps).to(device)
        self._i = 0

    def step(self, model_output: torch.Tensor, t, sample: torch.Tensor) -> torch.Tensor:
        i = self._i
        self._i += 1
        return sample + (self.sigmas[i + 1] - self.sigmas[i]).to(sample.dtype) * model_output


class LatentDecoder(nn.Module):
    """Small convolutional latent -> RGB decoder (x8 upsampling) standing in for the VAE decoder."""

    def __init__(self, in_channels: int = 4, width: int = 32) -> None:
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv2d(in_channels, width, 3, padding=1), nn.SiLU(),
            nn.Upsample(scale_factor=2), nn.Conv2d(width, width, 3, padding=1), nn.SiLU(),
            nn.Upsample(scale_factor=2), nn.Conv2d(width, width // 2, 3, padding=1), nn.SiLU(),
            nn.Upsample(scale_factor=2), nn.Conv2d(width // 2, 3, 3, padding=1))

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self.net(z.to(self.net[0].weight.dtype))


@dataclass
class DiffusionOutput:
    images: torch.Tensor
    latents: torch.Tensor


class DiffusionPipeline:
    """`pipe(prompt_embeds=[B, L, C] | prompt=..., num_inference_steps=N, guidance_scale=g)`; classifier-free
    guidance runs the conditional and unconditional branches as one batch of 2B."""

    def __init__(self, transformer: nn.Module, scheduler=None, vae: Optional[nn.Module] = None,
                 text_encoder: Optional[Callable] = None, tokenizer: Optional[Callable] = None) -> None:
        self.transformer = transformer
        cfg = transformer.cfg
        self.scheduler = scheduler or (FlowMatchEulerScheduler() if cfg.model_type == "stable_diffusion_3"
                                       else DDIMScheduler())
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer

    def to(self, device):
        self.transformer = self.transformer.to(device)
        if self.vae is not None:
            self.vae = self.vae.to(device)
        if isinstance(self.text_encoder, nn.Module):
            self.text_encoder = self.text_encoder.to(device)
        return self

    @property
    def device(self):
        return next(self.transformer.parameters()).device

    def encode_prompt(self, prompt: List[str]) -> torch.Tensor:
        assert self.text_encoder is not None and self.tokenizer is not None, \
            "pass prompt_embeds, or build the pipeline with a tokenizer + text encoder"
        ids = self.tokenizer(prompt)
        ids = ids["input_ids"] if isinstance(ids, dict) else ids
        ids = torch.as_tensor(ids, device=self.device)
        out = self.text_encoder(input_ids=ids)
        return out["last_hidden_state"] if isinstance(out, dict) else out

    @torch.no_grad()
    def __call__(self, prompt=None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None,
                 pooled_prompt_embeds: Optional[torch.Tensor] = None, num_inference_steps: int = 20,
                 guidance_scale: float = 4.5, generator: Optional[torch.Generator] = None,
                 latents: Optional[torch.Tensor] = None, output_type: str = "pt", **unused) -> DiffusionOutput:
        cfg = self.transformer.cfg
        dev = self.device
        dtype = next(self.transformer.parameters()).dtype
        if prompt_embeds is None:
            prompt_embeds = self.encode_prompt([prompt] if isinstance(prompt, str) else list(prompt))
        prompt_embeds = prompt_embeds.to(dev, dtype)
        B = prompt_embeds.shape[0]
        cfg_on = guidance_scale > 1.0
        if cfg_on:
            neg = negative_prompt_embeds.to(dev, dtype) if negative_prompt_embeds is not None \
                else torch.zeros_like(prompt_embeds)
            text = torch.cat([neg, prompt_embeds], 0)
        else:
            text = prompt_embeds
        pooled = None
        if cfg.model_type == "stable_diffusion_3":
            pooled = pooled_prompt_embeds.to(dev, dtype) if pooled_prompt_embeds is not None \
                else torch.zeros(B, cfg.pooled_projection_dim, device=dev, dtype=dtype)
            if cfg_on:
                pooled = torch.cat([torch.zeros_like(pooled), pooled], 0)
        if latents is None:
            latents = torch.randn(B, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=generator,
                                  device=generator.device if generator is not None else dev).to(dev, dtype)
        latents = latents * self.scheduler.init_noise_sigma
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        pp = getattr(self.transformer, "patch_parallel", None)
        if pp is not None:
            pp.reset()
        for t in self.scheduler.timesteps:
            inp = torch.cat([latents, latents], 0) if cfg_on else latents
            out = self.transformer(hidden_states=inp, encoder_hidden_states=text, timestep=t.reshape(1),
                                   pooled_projections=pooled)
            out = out[:, : cfg.in_channels]          # drop the learned-sigma half
            if cfg_on:
                un, co = out.chunk(2, 0)
                out = un + guidance_scale * (co - un)
            latents = self.scheduler.step(out, t, latents)
            if pp is not None:
                pp.next_step()
        images = self.vae.decode(latents) if self.vae is not None else latents
        return DiffusionOutput(images=images, latents=latents)


def build_diffusion_pipeline(name_or_config, with_decoder: bool = True) -> DiffusionPipeline:
    cfg = DIT_ZOO[name_or_config] if isinstance(name_or_config, str) else name_or_config
    model = (SD3Transformer2D if cfg.model_type == "stable_diffusion_3" else PixArtTransformer2D)(cfg)
    return DiffusionPipeline(model, vae=LatentDecoder(cfg.in_channels) if with_decoder else None)
