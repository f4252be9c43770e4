"""Native DiT backbones + samplers + Distrifusion patch parallelism (reference: tests/test_infer/test_models/
test_pixart_alpha.py / test_stablediffusion3.py and the distrifusion examples)."""
import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.inference.config import InferenceConfig
from colossalai_b200.inference.core.diffusion_engine import DiffusionEngine
from colossalai_b200.inference.modeling.layers import DistriConv2d, PatchParallelContext, enable_patch_parallel
from colossalai_b200.models.dit import DDIMScheduler, FlowMatchEulerScheduler, build_diffusion_pipeline
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


@pytest.mark.parametrize("name", ["pixart-tiny", "sd3-tiny"])
def test_pipeline_runs_and_is_deterministic(name):
    torch.manual_seed(0)
    pipe = build_diffusion_pipeline(name)
    emb = torch.randn(2, 6, 24)
    a = pipe(prompt_embeds=emb, num_inference_steps=3, guidance_scale=2.0, generator=torch.Generator().manual_seed(3))
    b = pipe(prompt_embeds=emb, num_inference_steps=3, guidance_scale=2.0, generator=torch.Generator().manual_seed(3))
    assert a.images.shape == (2, 3, 64, 64) and torch.equal(a.latents, b.latents)
    c = pipe(prompt_embeds=emb, num_inference_steps=3, guidance_scale=1.0, generator=torch.Generator().manual_seed(3))
    assert not torch.allclose(a.latents, c.latents)       # guidance changes the trajectory


def test_schedulers_reach_the_data_endpoint():
    x0 = torch.randn(1, 4, 8, 8)
    noise = torch.randn_like(x0)
    s = DDIMScheduler()
    s.set_timesteps(10)
    t0 = int(s.timesteps[0])
    a = s.alphas_cumprod[t0]
    x = a.sqrt() * x0 + (1 - a).sqrt() * noise
    for t in s.timesteps:               # an oracle that always predicts the true noise recovers x0 exactly
        a_t = s.alphas_cumprod[int(t)]
        eps = (x - a_t.sqrt() * x0) / (1 - a_t).sqrt()
        x = s.step(eps, t, x)
    torch.testing.assert_close(x, x0, atol=1e-4, rtol=1e-4)
    f = FlowMatchEulerScheduler()
    f.set_timesteps(8)
    x = noise.clone()
    for t in f.timesteps:               # rectified flow: v = noise - x0 along the straight path
        x = f.step(noise - x0, t, x)
    torch.testing.assert_close(x, x0, atol=1e-5, rtol=1e-5)


def test_engine_native_pipeline():
    eng = DiffusionEngine("pixart-tiny", InferenceConfig(dtype="fp32"))
    emb = torch.randn(1, 6, 24)
    imgs = eng.generate(prompts=emb, num_inference_steps=2, guidance_scale=1.0)
    assert len(imgs) == 1 and imgs[0].shape == (1, 3, 64, 64)


def _pp_worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    for name in ("pixart-tiny", "sd3-tiny"):
        torch.manual_seed(0)
        pipe = build_diffusion_pipeline(name, with_decoder=False)
        emb = torch.randn(2, 6, 24)
        gen = lambda: torch.Generator().manual_seed(5)
        ref = pipe(prompt_embeds=emb, num_inference_steps=4, guidance_scale=2.0, generator=gen()).latents
        ctx = enable_patch_parallel(pipe.transformer, None, mode="sync")
        got = pipe(prompt_embeds=emb, num_inference_steps=4, guidance_scale=2.0, generator=gen()).latents
        torch.testing.assert_close(got, ref, atol=1e-5, rtol=1e-5)     # synchronous exchange is exact
        ctx.mode, ctx.warmup_steps = "stale", 1
        stale = pipe(prompt_embeds=emb, num_inference_steps=4, guidance_scale=2.0, generator=gen()).latents
        assert torch.isfinite(stale).all() and ctx.step == 4
        assert (stale - ref).abs().max() < 0.5 * ref.abs().max()       # stale K/V: close, not exact
    # halo conv
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(3, 5, 3, padding=1)
    x = torch.randn(2, 3, 8, 6)
    ctx = PatchParallelContext(group=None, mode="sync")
    slab, _ = ctx.split_rows(x, 1)
    y = ctx.gather_rows(DistriConv2d(conv, ctx)(slab))
    torch.testing.assert_close(y, conv(x), atol=1e-5, rtol=1e-5)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_distrifusion_patch_parallel_world2():
    spawn(_pp_worker, 2)
