"""Generate with a (trained) sparse-MoE model under expert parallelism.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 applications/ColossalMoE/infer.py --model mixtral-tiny --ep 2 \
        --checkpoint ckpt/epoch-0_step-100 --prompt "The capital of France is"
    python applications/ColossalMoE/infer.py --model mixtral-tiny --engine        # one GPU / CPU: paged-KV engine

With `--ep N` every rank holds 1/N of the experts and all ranks decode the SAME batch in lock-step (the dispatch /
combine all-to-all needs every rank in every layer); greedy decoding keeps them identical, sampling uses a shared seed.
That path recomputes the prefix every step (no KV cache across the expert-parallel group) - it is the demonstration
the reference ships; for serving, a single rank (`--engine`) runs the continuous-batching `InferenceEngine` with paged
KV, CUDA graphs and the grouped expert GEMM.  Prompts are byte-tokenised unless `--tokenizer` names a HF tokenizer.

Parity: reference `applications/ColossalMoE/infer.py:1-110` (+ `infer.sh`).
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import colossalai_b200  # noqa: E402
from colossalai_b200.booster import Booster  # noqa: E402
from colossalai_b200.booster.plugin import MoeHybridParallelPlugin  # noqa: E402
from colossalai_b200.models import build_model, get_config  # noqa: E402
from colossalai_b200.testing import free_port  # noqa: E402
from utils import load_checkpoint  # noqa: E402


@torch.no_grad()
def decode(model, ids: torch.Tensor, max_new_tokens: int, temperature: float, eos: int, generator) -> torch.Tensor:
    done = torch.zeros(ids.shape[0], dtype=torch.bool, device=ids.device)
    for _ in range(max_new_tokens):
        logits = model(input_ids=ids)["logits"]                 # the zoo returns token-major logits [B * S, V]
        logits = logits.reshape(ids.shape[0], ids.shape[1], -1)[:, -1].float()
        if temperature > 0:
            probs = (logits / temperature).softmax(-1).cpu()
            nxt = torch.multinomial(probs, 1, generator=generator).squeeze(-1).to(ids.device)
        else:
            nxt = logits.argmax(-1)
        nxt = torch.where(done, torch.zeros_like(nxt), nxt)
        ids = torch.cat([ids, nxt[:, None]], 1)
        done |= nxt == eos
        if bool(done.all()):
            break
    return ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="mixtral-tiny")
    ap.add_argument("--pretrained", default=None, help="HuggingFace checkpoint directory")
    ap.add_argument("--checkpoint", default=None, help="directory written by train.py (epoch-E_step-S)")
    ap.add_argument("--ep", type=int, default=1)
    ap.add_argument("--engine", action="store_true", help="single rank: paged-KV InferenceEngine")
    ap.add_argument("--tokenizer", default=None)
    ap.add_argument("--prompt", nargs="*", default=["Hello, my name is", "The capital of France is"])
    ap.add_argument("--max_new_tokens", type=int, default=16)
    ap.add_argument("--temperature", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    if "RANK" in os.environ:
        colossalai_b200.launch_from_torch(backend="nccl" if torch.cuda.is_available() else "gloo")
    else:
        colossalai_b200.launch(0, 1, "127.0.0.1", free_port(), verbose=False)
    rank = dist.get_rank()
    dev = colossalai_b200.accelerator.get_accelerator().get_current_device()
    if args.tokenizer:
        from transformers import AutoTokenizer

        tok = AutoTokenizer.from_pretrained(args.tokenizer)
        enc = lambda t: tok(t, add_special_tokens=True)["input_ids"]            # noqa: E731
        dec = lambda ids: tok.decode(ids, skip_special_tokens=True)             # noqa: E731
        eos = tok.eos_token_id
    else:
        enc = lambda t: [1] + [3 + b for b in t.encode()]                        # noqa: E731
        dec = lambda ids: bytes(max(0, min(255, i - 3)) for i in ids if i >= 3).decode(errors="replace")   # noqa: E731
        eos = 2
    torch.manual_seed(42)
    if args.pretrained:
        from colossalai_b200.models.hf_io import load_hf_checkpoint

        model = load_hf_checkpoint(args.pretrained)
    else:
        model = build_model(get_config(args.model))
    prompts = [enc(p) for p in args.prompt]
    if args.engine:
        assert dist.get_world_size() == 1, "--engine is the single-rank serving path"
        from colossalai_b200.inference import InferenceConfig, InferenceEngine
        from colossalai_b200.inference.config import GenerationConfig

        if args.checkpoint:
            from colossalai_b200.checkpoint_io import GeneralCheckpointIO

            GeneralCheckpointIO().load_model(model, os.path.join(args.checkpoint, "modeling"))
        cuda = torch.cuda.is_available()
        model = model.eval() if cuda else model.float().eval()
        engine = InferenceEngine(model, None, InferenceConfig(max_batch_size=len(prompts), max_input_len=256,
                                                               max_output_len=args.max_new_tokens, block_size=16,
                                                               dtype="bf16" if cuda else "fp32"))
        _, outs = engine.generate(prompts_token_ids=prompts, return_token_ids=True,
                                  generation_config=GenerationConfig(max_new_tokens=args.max_new_tokens,
                                                                     do_sample=args.temperature > 0,
                                                                     temperature=max(args.temperature, 1e-5)))
        outs = [o[len(p):] for o, p in zip(outs, prompts)]      # the engine returns prompt + completion
    else:
        plugin = MoeHybridParallelPlugin(ep_size=args.ep, tp_size=1, pp_size=1, zero_stage=0,
                                         precision="bf16" if torch.cuda.is_available() else "fp32")
        booster = Booster(plugin=plugin)
        model, *_ = booster.boost(model)
        if args.checkpoint:
            load_checkpoint(args.checkpoint, booster, model)
        model.eval()
        width = max(len(p) for p in prompts)                    # left-pad with BOS so every row ends at its last token
        ids = torch.tensor([[1] * (width - len(p)) + p for p in prompts], device=dev)
        gen = torch.Generator().manual_seed(args.seed)         # the same draws on every rank
        out = decode(model, ids, args.max_new_tokens, args.temperature, eos, gen)
        outs = [row[width:].tolist() for row in out]
    if rank == 0:
        for p, o in zip(args.prompt, outs):
            o = o[: o.index(eos)] if eos in o else o
            print(f"[prompt] {p!r}\n[output] {dec(o)!r}  ({len(o)} tokens)")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
