"""Text-to-image through `DiffusionEngine`: PixArt-alpha (DDIM, classifier-free guidance) or Stable-Diffusion-3 (MM-DiT,
rectified-flow Euler) backbones, optionally with Distrifusion patch parallelism over the ranks.

    python examples/inference/stable_diffusion/sd3_generation.py --model sd3-tiny --steps 8 --out /tmp/sd3.pt
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/inference/stable_diffusion/sd3_generation.py \
        --model pixart-tiny --patch_parallel 2 --steps 8

`--model` is a name of `colossalai_b200.models.dit.DIT_ZOO` (`pixart-alpha-xl-2`, `sd3-medium`, or the `*-tiny`
shapes that run on a CPU) or, with the optional `diffusers` package, a pipeline directory.  The native pipelines take
pre-computed text embeddings (`--prompt_len` random vectors stand in for the T5 / CLIP encoders, which are not part of
the zoo); with a diffusers pipeline `--prompt` is encoded by the pipeline itself.  With `--patch_parallel N` every
rank denoises a horizontal slab of the latent and attends to the other slabs' keys / values from the previous step
(stale activations after `--warmup` synchronous steps); rank 0 holds the full image at the end.

Parity: reference `examples/inference/stable_diffusion/{sd3_generation.py, compute_metric.py, run_benchmark.sh}`.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))

import colossalai_b200  # noqa: E402
from colossalai_b200.inference.config import InferenceConfig  # noqa: E402
from colossalai_b200.inference.core.diffusion_engine import DiffusionEngine  # noqa: E402
from colossalai_b200.models.dit import DIT_ZOO  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="pixart-tiny")
    ap.add_argument("--prompt", default="a photo of an astronaut riding a horse")
    ap.add_argument("--prompt_len", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--guidance", type=float, default=4.5)
    ap.add_argument("--patch_parallel", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=2, help="synchronous steps before the stale exchange starts")
    ap.add_argument("--dtype", default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None, help="save the image tensor [B, 3, H, W] here (rank 0)")
    args = ap.parse_args()
    multi = "RANK" in os.environ
    if multi:
        colossalai_b200.launch_from_torch(backend="nccl" if torch.cuda.is_available() else "gloo")
    rank = dist.get_rank() if multi else 0
    dtype = args.dtype or ("bf16" if torch.cuda.is_available() else "fp32")
    cfg = InferenceConfig(dtype=dtype, patched_parallelism_size=args.patch_parallel)
    cfg.pp_warmup_steps = args.warmup
    torch.manual_seed(args.seed)                                       # same weights and noise on every rank
    engine = DiffusionEngine(args.model, cfg)
    kw = dict(num_inference_steps=args.steps, guidance_scale=args.guidance)
    if args.model in DIT_ZOO:
        tr = DIT_ZOO[args.model]
        prompts = torch.randn(args.batch, min(args.prompt_len, tr.max_text_len), tr.caption_channels,
                              generator=torch.Generator().manual_seed(args.seed))
        kw["generator"] = torch.Generator().manual_seed(args.seed + 1)
    else:
        prompts = [args.prompt] * args.batch
    t0 = time.perf_counter()
    images = engine.generate(prompts=prompts, **kw)[0]
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        img = images if torch.is_tensor(images) else images[0]
        shape = tuple(img.shape) if torch.is_tensor(img) else getattr(img, "size", None)
        print(f"generated {shape} in {dt:.2f} s ({args.steps} steps, guidance {args.guidance}, "
              f"patch parallel {args.patch_parallel}, {dtype})")
        if torch.is_tensor(img):
            print(f"pixel range [{float(img.min()):.3f}, {float(img.max()):.3f}], finite {bool(torch.isfinite(img).all())}")
            if args.out:
                torch.save(img.float().cpu(), args.out)
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
