"""Llama training throughput benchmark through the Booster API (reference: examples/language/llama/benchmark.py).

    torchrun --nproc-per-node 8 examples/language/llama/benchmark.py -c llama3-8b -p 3d --tp 8 --sp_mode split_gather \
        -b 8 -l 4096 --comm_backend fused

Plugins: 3d (HybridParallelPlugin), zero2 / zero1 (LowLevelZeroPlugin), gemini, ddp, fsdp.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

import colossalai_b200  # noqa: E402
from colossalai_b200.booster import Booster  # noqa: E402
from colossalai_b200.booster.plugin import (GeminiPlugin, HybridParallelPlugin, LowLevelZeroPlugin,  # noqa: E402
                                            TorchDDPPlugin, TorchFSDPPlugin)
from colossalai_b200.lazy import LazyInitContext  # noqa: E402
from colossalai_b200.models import build_model, get_config  # noqa: E402
from colossalai_b200.nn.optimizer import HybridAdam  # noqa: E402
from data_utils import RandomDataset, format_numel_str, get_model_numel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", default="llama-tiny", help="model zoo name (llama3-8b, llama2-7b, ...)")
    ap.add_argument("-p", "--plugin", default="3d", choices=["3d", "zero1", "zero2", "gemini", "ddp", "fsdp"])
    ap.add_argument("-b", "--batch_size", type=int, default=2, help="per-dp-rank batch size")
    ap.add_argument("-s", "--num_steps", type=int, default=5)
    ap.add_argument("-i", "--ignore_steps", type=int, default=2)
    ap.add_argument("-l", "--max_length", type=int, default=512)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--sp", type=int, default=1)
    ap.add_argument("--sp_mode", default=None, choices=[None, "split_gather", "ring", "all_to_all", "ring_attn"])
    ap.add_argument("--zero", type=int, default=0)
    ap.add_argument("--mbs", type=int, default=1, help="micro-batch size under pipeline parallelism")
    ap.add_argument("--pp_style", default="1f1b", choices=["1f1b", "interleaved", "zbv"])
    ap.add_argument("--n_chunks", type=int, default=1)
    ap.add_argument("-g", "--grad_checkpoint", action="store_true")
    ap.add_argument("--comm_backend", default="nccl", choices=["nccl", "fused"])
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--offload", action="store_true", help="gemini: keep optimizer states on the host")
    ap.add_argument("--offload_optim_frac", type=float, default=0.0,
                    help="3d plugin: fraction of the ZeRO optimizer state tiered to pinned host memory (needs --zero 1/2)")
    ap.add_argument("--layers", type=int, default=0, help="override the layer count (reduced-depth runs)")
    args = ap.parse_args()

    colossalai_b200.launch_from_torch()
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = get_config(args.config, **({"num_hidden_layers": args.layers} if args.layers else {}))
    if args.plugin == "3d":
        plugin = HybridParallelPlugin(tp_size=args.tp, pp_size=args.pp, sp_size=args.sp if args.sp > 1 else None,
                                      zero_stage=args.zero, precision=args.precision,
                                      enable_sequence_parallelism=args.sp_mode is not None,
                                      sequence_parallelism_mode=args.sp_mode,
                                      microbatch_size=None if args.pp_style == "zbv" else args.mbs,
                                      num_microbatches=(args.batch_size // args.mbs) if args.pp_style == "zbv" else None,
                                      pp_style=args.pp_style, num_model_chunks=args.n_chunks, max_norm=1.0,
                                      comm_backend=args.comm_backend,
                                      cpu_offload=args.offload_optim_frac > 0,
                                      offload_optim_frac=args.offload_optim_frac if args.offload_optim_frac > 0 else 1.0)
    elif args.plugin in ("zero1", "zero2"):
        plugin = LowLevelZeroPlugin(stage=int(args.plugin[-1]), precision=args.precision, max_norm=1.0)
    elif args.plugin == "gemini":
        plugin = GeminiPlugin(precision=args.precision, placement_policy="static",
                              offload_optim_frac=1.0 if args.offload else 0.0, max_norm=1.0)
    elif args.plugin == "ddp":
        plugin = TorchDDPPlugin()
    else:
        plugin = TorchFSDPPlugin()
    booster = Booster(plugin=plugin)
    dp_size = getattr(plugin, "dp_size", world)
    dataset = RandomDataset(num_samples=args.batch_size * args.num_steps * dp_size, max_length=args.max_length,
                            vocab_size=cfg.vocab_size)
    loader = plugin.prepare_dataloader(dataset, batch_size=args.batch_size, shuffle=True, drop_last=True,
                                       pin_memory=torch.cuda.is_available())
    with LazyInitContext():
        model = build_model(cfg)
    if args.grad_checkpoint:
        model.gradient_checkpointing_enable()
    optimizer = HybridAdam(model.parameters(), lr=1e-5, weight_decay=0.1)
    model, optimizer, _, loader, _ = booster.boost(model, optimizer, dataloader=loader)
    if rank == 0:
        print(f"model {args.config}: {format_numel_str(get_model_numel(model))} params on this rank, plugin {args.plugin}")
    dev = colossalai_b200.accelerator.get_accelerator().get_current_device()
    use_events = dev.type == "cuda"
    times = []
    it = iter(loader)
    for step in range(args.num_steps):
        t0 = time.perf_counter()
        if use_events:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        if args.plugin == "3d" and args.pp > 1:
            out = booster.execute_pipeline(it, model, lambda o, b: o["loss"], optimizer, return_loss=True)
            loss = out["loss"]
        else:
            batch = {k: v.to(dev, non_blocking=True) for k, v in next(it).items()}
            loss = model(input_ids=batch["input_ids"], labels=batch["labels"])["loss"]
            booster.backward(loss, optimizer)
        optimizer.step()
        optimizer.zero_grad()
        if use_events:
            ev1.record()
            torch.cuda.synchronize()
            dt = ev0.elapsed_time(ev1) / 1e3
        else:
            dt = time.perf_counter() - t0
        if step >= args.ignore_steps:
            times.append(dt)
        if args.plugin == "3d" and args.pp > 1:
            # the loss lives on the last pipeline stage: share it for the log line
            lt = torch.full((1,), float("-inf"), device=dev) if loss is None else loss.detach().float().reshape(1).to(dev)
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
            loss = lt[0]
        if rank == 0:
            print(f"step {step}: loss {float(loss) if loss is not None else float('nan'):.4f}  {dt * 1e3:.1f} ms")
    if times:
        t = torch.tensor(sum(times) / len(times), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tok = args.batch_size * dp_size * args.max_length / t.item()
        if rank == 0:
            print(f"throughput: {tok:,.0f} tokens/s (max over ranks), {t.item() * 1e3:.1f} ms/step")
        if use_events:
            peak = torch.tensor(torch.cuda.max_memory_allocated() / 2**30, device=dev)
            dist.all_reduce(peak, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"peak device memory: {peak.item():.1f} GiB (max over ranks)")
    dist.barrier()
    colossalai_b200.initialize.shutdown()


if __name__ == "__main__":
    main()
