"""Continual pre-training / supervised fine-tuning driver with resumable checkpoints.

    # synthetic documents, ZeRO-2, checkpoints every 20 steps
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 applications/Colossal-LLaMA/train.py \
        --model llama-tiny --plugin zero2 --max_length 128 --steps 40 --save_dir /tmp/cl --save_interval 20
    # resume (same command + --load_checkpoint): picks up model, optimizer, LR schedule and the data position
    torchrun ... applications/Colossal-LLaMA/train.py ... --load_checkpoint /tmp/cl/epoch-0_step-20

`--dataset` takes jsonl files of `{"source": ..., "target": ...}` records (`source` carries no loss); they are tokenised
(byte-level unless `--tokenizer` names a HuggingFace tokenizer directory) and SPLICED into constant-length sequences, so
no step is spent on padding.  `--pretrained` is a HuggingFace checkpoint directory imported into the native zoo
(`models.hf_io`); `--expand_vocab N` grows the embedding / head for continual pre-training on a new language,
`--freeze_non_embeds` trains only the new rows' neighbourhood (embeddings + head), `--neftune` adds embedding noise
for SFT.  Plugins: `ddp`, `zero1`, `zero2`, `gemini`, `3d` (`--tp/--pp/--sp_mode`).

Parity: reference `applications/Colossal-LLaMA/train.py:1-539` (+ `train.example.sh`, `train_sft.example.sh`).
"""
import argparse
import json
import math
import os
import sys
import time

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import colossalai_b200  # noqa: E402
from colossal_llama import (ClosedToConstantLengthSplicedDataset, activate_neftune, expand_vocab,  # noqa: E402
                            format_numel_str, freeze_non_embeds_parameters, get_model_numel, load_checkpoint,
                            save_checkpoint, supervised_tokenize_pretrain)
from colossalai_b200.booster import Booster  # noqa: E402
from colossalai_b200.booster.plugin import (GeminiPlugin, HybridParallelPlugin, LowLevelZeroPlugin,  # noqa: E402
                                            TorchDDPPlugin)
from colossalai_b200.cluster import DistCoordinator  # noqa: E402
from colossalai_b200.models import build_model, get_config  # noqa: E402
from colossalai_b200.nn.lr_scheduler import CosineAnnealingWarmupLR  # noqa: E402
from colossalai_b200.nn.optimizer import HybridAdam  # noqa: E402
from colossalai_b200.testing import free_port  # noqa: E402


def byte_tokenizer(text: str):
    return [3 + b for b in text.encode()]            # 0 pad / 1 bos / 2 eos, bytes at 3..258


def documents(args, vocab: int):
    if args.dataset:
        for path in args.dataset:
            with open(path) as f:
                for line in f:
                    if line.strip():
                        yield json.loads(line)
        return
    g = torch.Generator().manual_seed(1234)           # synthetic corpus: arithmetic progressions spelled in bytes
    for _ in range(args.synthetic_docs):
        a, d = int(torch.randint(0, 50, (1,), generator=g)), int(torch.randint(1, 9, (1,), generator=g))
        n = int(torch.randint(8, 40, (1,), generator=g))
        yield {"source": f"start {a} step {d}: ", "target": " ".join(str(a + i * d) for i in range(n))}


def build_plugin(args, precision: str):
    if args.plugin == "ddp":
        return TorchDDPPlugin()
    if args.plugin in ("zero1", "zero2"):
        return LowLevelZeroPlugin(stage=int(args.plugin[-1]), precision=precision, max_norm=args.grad_clip)
    if args.plugin == "gemini":
        return GeminiPlugin(precision=precision, placement_policy="static", max_norm=args.grad_clip)
    sp = dict(enable_sequence_parallelism=True, sequence_parallelism_mode=args.sp_mode) if args.sp_mode else {}
    return HybridParallelPlugin(tp_size=args.tp, pp_size=args.pp, zero_stage=args.zero_stage, precision=precision,
                                max_norm=args.grad_clip, num_microbatches=args.microbatches if args.pp > 1 else None, **sp)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-tiny", help="zoo preset (architecture)")
    ap.add_argument("--pretrained", default=None, help="HuggingFace checkpoint directory to start from")
    ap.add_argument("--dataset", nargs="*", default=None, help="jsonl files of {source, target}")
    ap.add_argument("--tokenizer", default=None)
    ap.add_argument("--synthetic_docs", type=int, default=2000)
    ap.add_argument("--plugin", default="zero2", choices=["ddp", "zero1", "zero2", "gemini", "3d"])
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--zero_stage", type=int, default=0)
    ap.add_argument("--sp_mode", default=None)
    ap.add_argument("--microbatches", type=int, default=2)
    ap.add_argument("--precision", default=None)
    ap.add_argument("--max_length", type=int, default=256)
    ap.add_argument("--batch_size", type=int, default=4, help="per data-parallel rank")
    ap.add_argument("--accumulation_steps", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40, help="optimizer steps")
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--warmup_steps", type=int, default=None)
    ap.add_argument("--weight_decay", type=float, default=0.1)
    ap.add_argument("--grad_clip", type=float, default=1.0)
    ap.add_argument("--expand_vocab", type=int, default=0, help="new vocabulary size (continual pre-training)")
    ap.add_argument("--freeze_non_embeds", action="store_true")
    ap.add_argument("--neftune", type=float, default=0.0, help="NEFTune noise alpha (SFT)")
    ap.add_argument("--save_dir", default=None)
    ap.add_argument("--save_interval", type=int, default=0)
    ap.add_argument("--load_checkpoint", default=None)
    ap.add_argument("--log_file", default=None, help="jsonl of per-step metrics (rank 0)")
    args = ap.parse_args()
    if "RANK" in os.environ:
        colossalai_b200.launch_from_torch(backend="nccl" if torch.cuda.is_available() else "gloo")
    else:
        colossalai_b200.launch(0, 1, "127.0.0.1", free_port(), verbose=False)
    coordinator = DistCoordinator()
    cuda = torch.cuda.is_available()
    dev = colossalai_b200.accelerator.get_accelerator().get_current_device()
    precision = args.precision or ("bf16" if cuda or args.plugin != "ddp" else "fp32")

    # ---- model
    torch.manual_seed(42)
    if args.pretrained:
        from colossalai_b200.models.hf_io import load_hf_checkpoint

        model = load_hf_checkpoint(args.pretrained, dtype=torch.float32)
    else:
        model = build_model(get_config(args.model))
    if args.expand_vocab:
        model = expand_vocab(model, args.expand_vocab)
    if args.freeze_non_embeds:
        freeze_non_embeds_parameters(model)
    if args.neftune > 0:
        activate_neftune(model, args.neftune)
    cfg = model.cfg
    coordinator.print_on_master(f"model {cfg.model_type}: {format_numel_str(get_model_numel(model))} parameters, "
                                f"{format_numel_str(get_model_numel(model, trainable_only=True))} trainable")

    # ---- data: tokenise -> splice to constant length -> shard over the data-parallel ranks
    if args.tokenizer:
        from transformers import AutoTokenizer

        hf_tok = AutoTokenizer.from_pretrained(args.tokenizer)
        tok = lambda t: hf_tok(t, add_special_tokens=False)["input_ids"]      # noqa: E731
    else:
        tok = byte_tokenizer
    tokenised = (supervised_tokenize_pretrain(d, tok, max_length=args.max_length) for d in documents(args, cfg.vocab_size))
    spliced = list(ClosedToConstantLengthSplicedDataset(list(tokenised), max_length=args.max_length, num_packed_sequences=16))
    assert all(int(s["input_ids"].max()) < cfg.vocab_size for s in spliced), "token id outside the model's vocabulary"

    optimizer = HybridAdam([p for p in model.parameters() if p.requires_grad], lr=args.lr, weight_decay=args.weight_decay,
                           betas=(0.9, 0.95))
    warmup = args.warmup_steps if args.warmup_steps is not None else max(1, int(0.025 * args.steps))
    scheduler = CosineAnnealingWarmupLR(optimizer, total_steps=args.steps, warmup_steps=warmup, eta_min=0.1 * args.lr)
    plugin = build_plugin(args, precision)
    booster = Booster(plugin=plugin)
    model, optimizer, _, _, scheduler = booster.boost(model, optimizer, lr_scheduler=scheduler)
    mesh = getattr(plugin, "pg_mesh", None)              # 3d plugin: data-parallel coordinates inside the mesh
    dp_size = plugin.dp_size if mesh is not None else dist.get_world_size()
    dp_rank = mesh.axis_rank("dp") if mesh is not None else dist.get_rank()

    start_step, sample_start = 0, 0
    if args.load_checkpoint:
        _, start_step, sample_start = load_checkpoint(args.load_checkpoint, booster, model, optimizer, scheduler)
        coordinator.print_on_master(f"resumed from {args.load_checkpoint}: step {start_step}, sample {sample_start}")

    per_step = args.batch_size * dp_size * args.accumulation_steps     # samples consumed per optimizer step
    log = open(args.log_file, "a") if args.log_file and coordinator.is_master() else None
    use_pp = args.plugin == "3d" and args.pp > 1
    t_last = time.perf_counter()
    for step in range(start_step, args.steps):
        total = torch.zeros((), device=dev)
        for micro in range(args.accumulation_steps):
            base = sample_start + (step - start_step) * per_step + micro * args.batch_size * dp_size + dp_rank * args.batch_size
            rows = [spliced[(base + i) % len(spliced)] for i in range(args.batch_size)]
            batch = {k: torch.stack([r[k] for r in rows]).to(dev) for k in ("input_ids", "labels")}
            if use_pp:
                out = booster.execute_pipeline(iter([batch]), model, lambda o, b: o["loss"], optimizer, return_loss=True)
                loss = out["loss"]
            else:
                loss = model(**batch)["loss"] / args.accumulation_steps
                booster.backward(loss, optimizer)
            if loss is not None:
                total += loss.detach().float()
        optimizer.step()
        optimizer.zero_grad()
        scheduler.step()
        dist.all_reduce(total, op=dist.ReduceOp.MAX if use_pp else dist.ReduceOp.SUM)
        mean = float(total) if use_pp else float(total) / dist.get_world_size()
        if coordinator.is_master():
            now = time.perf_counter()
            toks = per_step * args.max_length / max(now - t_last, 1e-9)
            t_last = now
            rec = {"step": step + 1, "loss": round(mean, 4), "ppl": round(math.exp(min(mean, 20.0)), 2),
                   "lr": optimizer.param_groups[0]["lr"], "tokens_per_s": round(toks)}
            print(f"step {rec['step']:4d} loss {rec['loss']:.4f} ppl {rec['ppl']:.2f} lr {rec['lr']:.3g} "
                  f"{rec['tokens_per_s']} tok/s")
            if log is not None:
                log.write(json.dumps(rec) + "\n")
                log.flush()
        if args.save_dir and args.save_interval and (step + 1) % args.save_interval == 0:
            path = save_checkpoint(args.save_dir, booster, model, optimizer, scheduler, 0, step + 1, per_step, coordinator,
                                   sampler_start_idx=sample_start + (step + 1 - start_step) * per_step)
            coordinator.print_on_master(f"saved {path}")
    if log is not None:
        log.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
