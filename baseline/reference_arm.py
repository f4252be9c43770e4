"""Reference arm of bench.py: runs the UNMODIFIED reference (hpcaitech/ColossalAI 0.5.0 installed into baseline/_ref with
`pip install --no-deps --target baseline/_ref`) through its own public API (`colossalai.launch_from_torch`, `Booster`,
a stock plugin) on the same metric/config as our arm: Llama-3-8B shape, bf16, seq 4096, one sequence per GPU per step,
AdamW, synthetic tokens, random init; K device-timed steps, max over ranks.

What is and is not possible offline (recorded in DESIGN.md):
  * the reference hard-imports packages that are not in this image (peft, galore_torch, bitsandbytes, ...).  They are
    never exercised by the benchmark path, so this file registers EMPTY stub modules for them before importing the
    reference — the reference's own files are untouched;
  * the reference's Shardformer policies import transformers-4.51 internals (`StaticCache` from modeling_llama ...) that
    do not exist in the installed transformers 5.5, so its TP/SP/PP path (HybridParallelPlugin) cannot be imported here.
    The arm therefore tries, in order, HybridParallelPlugin(tp=N) -> LowLevelZeroPlugin(stage=2) -> TorchDDPPlugin and
    reports which stock plugin actually ran in `config.parallelism`.
"""
from __future__ import annotations

import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
STUBS = ["peft", "galore_torch", "bitsandbytes", "ray", "rpyc", "diffusers", "tensornvme", "apex", "xformers"]


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder:
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in STUBS:
            from importlib.machinery import ModuleSpec

            return ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _unavailable(why: str) -> None:
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"impl": "reference", "unavailable": why.replace("\n", " ")[:300]}), flush=True)


def run_reference_arm(args) -> None:
    if not os.path.isdir(os.path.join(REF, "colossalai")):
        _unavailable("baseline/_ref is missing (pip install --no-deps --target baseline/_ref /root/reference not done)")
        return
    # make sure OUR package is not importable by accident on this path
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    sys.path.insert(0, REF)
    for name in STUBS:
        try:
            __import__(name)
        except Exception:
            if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
                sys.meta_path.append(_StubFinder())
    try:
        import torch
        import torch.distributed as dist
        import colossalai
        from colossalai.booster import Booster
    except Exception as e:
        _unavailable(f"reference import failed: {type(e).__name__}: {e}")
        return
    if not torch.cuda.is_available():
        _unavailable("no CUDA device")
        return

    if "RANK" not in os.environ:
        os.environ.update(RANK="0", LOCAL_RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1",
                          MASTER_PORT=str(29500 + os.getpid() % 2000))
    try:
        colossalai.launch_from_torch()
    except Exception as e:
        _unavailable(f"colossalai.launch_from_torch failed: {type(e).__name__}: {e}")
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    from transformers import LlamaConfig, LlamaForCausalLM

    shapes = {"llama3-8b": dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                                num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192,
                                rope_theta=500000.0)}
    hf_cfg = LlamaConfig(**shapes.get(args.model, shapes["llama3-8b"]))
    if args.layers:
        hf_cfg.num_hidden_layers = args.layers
    attn_impl = "sdpa"
    try:
        import flash_attn  # noqa: F401

        attn_impl = "flash_attention_2"
    except Exception:
        pass

    def build_model():
        torch.manual_seed(1234)
        with torch.device(dev):
            try:
                m = LlamaForCausalLM._from_config(hf_cfg, attn_implementation=attn_impl, torch_dtype=torch.bfloat16)
            except Exception:
                m = LlamaForCausalLM._from_config(hf_cfg, attn_implementation="sdpa", torch_dtype=torch.bfloat16)
        return m

    # transformers-5.x moved the cache classes out of modeling_llama; the reference (written against 4.51) imports them
    # from there.  Re-export them in place (environment shim — no reference file is touched).
    try:
        import transformers.cache_utils as _cu
        import transformers.models.llama.modeling_llama as _ml

        for _n in ("StaticCache", "DynamicCache", "Cache"):
            if not hasattr(_ml, _n) and hasattr(_cu, _n):
                setattr(_ml, _n, getattr(_cu, _n))
    except Exception:
        pass

    import gc
    import traceback

    chosen = None
    model = optimizer = booster = None
    errors = []
    plans = [("hybrid_tp", False), ("hybrid_tp", True), ("zero2", False), ("zero2", True), ("ddp", False)]
    for kind, ckpt in plans:
        label = {"hybrid_tp": f"tp{world}" + ("+sp(split_gather)" if world > 1 else ""), "zero2": f"zero2(dp{world})",
                 "ddp": f"ddp(dp{world})"}[kind] + ("+grad_ckpt" if ckpt else "")
        try:
            if kind == "hybrid_tp":
                from colossalai.booster.plugin import HybridParallelPlugin

                plugin = HybridParallelPlugin(tp_size=world, pp_size=1, precision="bf16", zero_stage=0,
                                              enable_sequence_parallelism=world > 1,
                                              sequence_parallelism_mode="split_gather" if world > 1 else None,
                                              enable_flash_attention=True, enable_fused_normalization=False,
                                              max_norm=1.0)
            elif kind == "zero2":
                from colossalai.booster.plugin import LowLevelZeroPlugin

                plugin = LowLevelZeroPlugin(stage=2, precision="bf16", max_norm=1.0)
            else:
                from colossalai.booster.plugin import TorchDDPPlugin

                plugin = TorchDDPPlugin()
            booster = Booster(plugin=plugin)
            model = build_model()
            if ckpt:
                model.gradient_checkpointing_enable()
            model.train()
            optimizer = torch.optim.AdamW(model.parameters(), lr=1e-5, betas=(0.9, 0.95), weight_decay=0.1)
            model, optimizer, _, _, _ = booster.boost(model, optimizer)
            # one probing step
            nseq = args.mbs * (world if kind == "hybrid_tp" else 1)
            ids = torch.randint(0, hf_cfg.vocab_size, (nseq, args.seq), device=dev)
            out = model(input_ids=ids, labels=ids)
            booster.backward(out.loss, optimizer)
            optimizer.step()
            optimizer.zero_grad()
            torch.cuda.synchronize()
            del out, ids
            chosen = (kind, label)
            break
        except Exception as e:  # try the next stock configuration
            msg = f"{label}: {type(e).__name__}: {str(e)[:200]}"
            errors.append(msg)
            if rank == 0:
                print("[reference arm] " + msg, file=sys.stderr, flush=True)
            traceback.clear_frames(e.__traceback__)
            e = None
            model = optimizer = booster = plugin = None
            gc.collect()
            torch.cuda.empty_cache()
            try:
                torch.cuda.reset_peak_memory_stats()
            except Exception:
                pass
    if chosen is None:
        _unavailable("no stock plugin of the reference runs on this stack: " + " | ".join(e_[:120] for e_ in errors))
        return
    tp_mode = chosen[0] == "hybrid_tp"

    # TP: every rank feeds the same global batch of mbs*world sequences (as our arm does); DP: mbs sequences per rank.
    S = args.seq
    B = args.mbs * world if tp_mode else args.mbs
    tokens_per_step = args.mbs * world * S * args.accum
    gen = torch.Generator().manual_seed(4321 + (0 if tp_mode else rank))
    host_ids = [torch.randint(0, hf_cfg.vocab_size, (args.accum, B, S), generator=gen).pin_memory() for _ in range(4)]
    dev_ids = [h.to(dev) for h in host_ids]

    def step(ids_dev):
        loss_acc = 0.0
        for a in range(args.accum):
            out = model(input_ids=ids_dev[a], labels=ids_dev[a])
            loss = out.loss / args.accum
            booster.backward(loss, optimizer)
            loss_acc = loss.detach() + loss_acc
        optimizer.step()
        optimizer.zero_grad()
        return loss_acc

    def timed(n, e2e):
        dist.barrier()
        torch.cuda.synchronize()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s_ev.record()
        last = None
        for i in range(n):
            if e2e:
                last = step(host_ids[i % 4].to(dev, non_blocking=True)).item()
            else:
                last = step(dev_ids[i % 4])
        e_ev.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([s_ev.elapsed_time(e_ev), (time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), (last.item() if torch.is_tensor(last) else last)

    sys.path.insert(0, ROOT)            # only for the shared nvidia-smi sampler helper living in bench.py
    try:
        from bench import ClockSampler
    except Exception:
        ClockSampler = None
    for i in range(max(args.warmup - 1, 0)):
        step(dev_ids[i % 4])
    sampler = ClockSampler(torch.cuda.current_device()) if (ClockSampler and rank == 0) else None
    if sampler:
        sampler.start()
    ms, wall, loss_val = timed(args.steps, False)
    clocks = sampler.stop() if sampler else {}
    e2e = None
    if not args.no_e2e:
        ms_e, wall_e, _ = timed(args.steps, True)
        e2e = {"value": tokens_per_step * args.steps / (max(ms_e, wall_e) / 1e3), "unit": "tokens/s",
               "h2d_bytes_per_step": host_ids[0].numel() * 8, "d2h_bytes_per_step": 4}
    value = tokens_per_step * args.steps / (ms / 1e3)
    res = {"metric": "tokens/sec (whole job, device-timed, max over ranks) Llama-3-8B training step (fwd+bwd+AdamW)",
           "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic tokens, random-init weights", "impl": "reference",
           "config": {"model": args.model, "global_batch": args.mbs * world * args.accum, "seq_len": S,
                      "parallelism": chosen[1], "reference_plugin": chosen[0], "attn": attn_impl,
                      "optimizer": "torch.optim.AdamW inside the reference's OptimizerWrapper",
                      "skipped_plugins": errors},
           "clocks": clocks, "gpu_launches": 0, "loss": loss_val,
           "peak_mem_mib": torch.cuda.max_memory_allocated() / 2**20}
    if e2e:
        res["e2e"] = e2e
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
