"""One fine-tuning driver for every model family in the zoo (decoder LMs, BERT, T5, Whisper, ViT, BLIP-2) and every
booster plugin — the per-family example directories call into it with their defaults.

    torchrun --nproc-per-node 2 examples/finetune.py --model gpt2-tiny --plugin hybrid --tp 2
    python examples/finetune.py --model t5-tiny --plugin zero2 --steps 10
    python examples/finetune.py --model vit-tiny --plugin gemini

Data is synthetic (random tokens / pixels / mel frames) with a learnable structure (the label is a function of the
input) so the loss visibly goes down; swap `make_batch` for a real dataloader.

Parity: reference `examples/language/{gpt,bert,opt,llama,deepseek,...}/`, `examples/images/vit`, each a
`Booster` + plugin + HF model training script.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import colossalai_b200  # noqa: E402
from colossalai_b200.booster import Booster  # noqa: E402
from colossalai_b200.booster.plugin import (GeminiPlugin, HybridParallelPlugin, LowLevelZeroPlugin,  # noqa: E402
                                            TorchDDPPlugin)
from colossalai_b200.models import build_model  # noqa: E402
from colossalai_b200.nn.lr_scheduler import CosineAnnealingWarmupLR  # noqa: E402
from colossalai_b200.nn.optimizer import HybridAdam  # noqa: E402


def make_batch(model_type: str, cfg, batch: int, seq: int, device, gen: torch.Generator):
    if model_type == "vit":
        x = torch.randn(batch, cfg.num_channels, cfg.image_size, cfg.image_size, generator=gen)
        y = (x.mean((1, 2, 3)) > 0).long() % cfg.num_labels
        return dict(pixel_values=x.to(device), labels=y.to(device))
    if model_type == "sam":
        x = torch.randn(batch, 3, cfg.image_size, cfg.image_size, generator=gen)
        pts = torch.rand(batch, 2, 2, generator=gen) * cfg.image_size
        side = 4 * cfg.grid
        yy, xx = torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij")
        centre = pts[:, 0] * side / cfg.image_size          # target: a disc around the first prompt point
        mask = (((xx[None] - centre[:, 0, None, None]) ** 2 + (yy[None] - centre[:, 1, None, None]) ** 2) < (side / 4) ** 2)
        mask = mask.float()[:, None].expand(-1, cfg.num_multimask_outputs, -1, -1).contiguous()
        return dict(pixel_values=x.to(device), input_points=pts.to(device),
                    input_labels=torch.ones(batch, 2, dtype=torch.long, device=device), labels=mask.to(device))
    if model_type == "t5":
        ids = torch.randint(2, cfg.vocab_size, (batch, seq), generator=gen)
        return dict(input_ids=ids.to(device), labels=ids[:, : seq // 2].flip(1).contiguous().to(device))
    if model_type == "whisper":
        feats = torch.randn(batch, cfg.num_mel_bins, 2 * min(seq, cfg.max_source_positions), generator=gen)
        lab = torch.randint(3, cfg.vocab_size, (batch, min(seq // 2, cfg.max_target_positions)), generator=gen)
        return dict(input_features=feats.to(device), labels=lab.to(device))
    if model_type == "blip2":
        ids = torch.randint(3, cfg.vocab_size, (batch, seq), generator=gen)
        x = torch.randn(batch, 3, cfg.image_size, cfg.image_size, generator=gen)
        return dict(pixel_values=x.to(device), input_ids=ids.to(device), labels=ids.to(device))
    # decoder / encoder LMs: learnable periodic token stream
    start = torch.randint(0, cfg.vocab_size, (batch, 1), generator=gen)
    ids = (start + torch.arange(seq)[None] * 3) % cfg.vocab_size
    return dict(input_ids=ids.to(device), labels=ids.to(device))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-tiny")
    ap.add_argument("--plugin", default="zero2", choices=["ddp", "zero1", "zero2", "gemini", "hybrid", "moe_hybrid"])
    ap.add_argument("--ep", type=int, default=1, help="expert-parallel size (moe_hybrid)")
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--sp_mode", default=None)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--seq", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--save", default=None, help="directory for a sharded checkpoint at the end")
    args = ap.parse_args()
    if "RANK" in os.environ:
        colossalai_b200.launch_from_torch(backend="nccl" if torch.cuda.is_available() else "gloo")
    else:
        colossalai_b200.launch(0, 1, "127.0.0.1", 29500 + os.getpid() % 1000,
                               backend="nccl" if torch.cuda.is_available() else "gloo")
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    precision = args.precision if dev.type == "cuda" else "fp32"
    if args.plugin == "ddp":
        plugin = TorchDDPPlugin()
    elif args.plugin in ("zero1", "zero2"):
        plugin = LowLevelZeroPlugin(stage=int(args.plugin[-1]), precision=precision, max_norm=1.0)
    elif args.plugin == "gemini":
        plugin = GeminiPlugin(precision=precision if precision != "fp32" else "bf16", max_norm=1.0,
                              **(dict(min_chunk_size_m=1, search_range_m=1) if dev.type == "cpu" else {}))
    elif args.plugin == "moe_hybrid":
        from colossalai_b200.booster.plugin import MoeHybridParallelPlugin

        plugin = MoeHybridParallelPlugin(tp_size=args.tp, pp_size=args.pp, ep_size=args.ep, precision=precision,
                                         max_norm=1.0, num_microbatches=args.batch if args.pp > 1 else None)
    else:
        plugin = HybridParallelPlugin(tp_size=args.tp, pp_size=args.pp, precision=precision, max_norm=1.0,
                                      enable_sequence_parallelism=args.sp_mode is not None,
                                      sequence_parallelism_mode=args.sp_mode,
                                      num_microbatches=args.batch if args.pp > 1 else None)
    booster = Booster(plugin=plugin)
    torch.manual_seed(0)
    model = build_model(args.model)
    cfg = model.cfg
    mt = getattr(cfg, "model_type", "llama")
    optimizer = HybridAdam(model.parameters(), lr=args.lr, weight_decay=0.01)
    sched = CosineAnnealingWarmupLR(optimizer, total_steps=args.steps, warmup_steps=max(1, args.steps // 10))
    model, optimizer, _, _, sched = booster.boost(model, optimizer, lr_scheduler=sched)
    gen = torch.Generator().manual_seed(1234 + (dist.get_rank() if args.plugin not in ("hybrid", "moe_hybrid") else 0))
    first = last = None
    t0 = time.perf_counter()
    for step in range(args.steps):
        batch = make_batch(mt, cfg, args.batch, args.seq, dev, gen)
        if args.plugin in ("hybrid", "moe_hybrid") and args.pp > 1:
            out = booster.execute_pipeline(iter([batch]), model, lambda o, b: o["loss"], optimizer, return_loss=True)
            loss = out["loss"]
        else:
            loss = model(**batch)["loss"]
            booster.backward(loss, optimizer)
        optimizer.step()
        optimizer.zero_grad()
        sched.step()
        if loss is not None:
            last = float(loss.detach())
            first = last if first is None else first
            if dist.get_rank() == 0 and (step % max(1, args.steps // 10) == 0 or step == args.steps - 1):
                print(f"step {step:4d} loss {last:.4f}", flush=True)
    if dist.get_rank() == 0:
        print(f"done: loss {first:.4f} -> {last:.4f} in {time.perf_counter() - t0:.1f}s "
              f"({args.model}, plugin={args.plugin}, world={dist.get_world_size()})")
    if args.save:
        booster.save_model(model, args.save, shard=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
