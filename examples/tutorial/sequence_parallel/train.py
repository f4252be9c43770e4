"""Sequence parallelism on an encoder (BERT) and a decoder (Llama): the same script, four ways to split the sequence.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/tutorial/sequence_parallel/train.py --mode split_gather
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/tutorial/sequence_parallel/train.py --mode all_to_all
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/tutorial/sequence_parallel/train.py --mode ring_attn --model llama-tiny

| mode | group | what moves |
|---|---|---|
| `split_gather` | the TP group | Megatron SP: activations outside attention / MLP are sequence shards; all-gather in front of the column linears, reduce-scatter behind the row linears (fused into the GEMM kernels on NVLink) |
| `ring` | the TP group | same layout, the gather / scatter decomposed into ring hops overlapped with partial GEMMs |
| `all_to_all` | its own SP group | Ulysses: sequence shards everywhere, one all-to-all swaps sequence <-> heads around attention |
| `ring_attn` | its own SP group | zigzag ring attention: K/V blocks travel round the ring, softmax state merged per hop (causal models) |

Every mode prints the same loss curve as the single-process run of the same seed (`--mode none`) - that is the point of
the tutorial.  Parity: reference `examples/tutorial/sequence_parallel/train.py` (BERT with its legacy ring self-attention)
and the SP modes of `HybridParallelPlugin` (`docs/source/en/features/sequence_parallelism.md`).
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))

import colossalai_b200  # noqa: E402
from colossalai_b200.booster import Booster  # noqa: E402
from colossalai_b200.booster.plugin import HybridParallelPlugin  # noqa: E402
from colossalai_b200.models import build_model, get_config  # noqa: E402
from colossalai_b200.nn.optimizer import HybridAdam  # noqa: E402
from colossalai_b200.testing import free_port  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="split_gather", choices=["none", "split_gather", "ring", "all_to_all", "ring_attn"])
    ap.add_argument("--model", default="bert-tiny")
    ap.add_argument("--seq", type=int, default=64)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--lr", type=float, default=2e-3)
    args = ap.parse_args()
    if "RANK" in os.environ:
        colossalai_b200.launch_from_torch(backend="nccl" if torch.cuda.is_available() else "gloo")
    else:
        colossalai_b200.launch(0, 1, "127.0.0.1", free_port(), verbose=False)
    world, rank = dist.get_world_size(), dist.get_rank()
    cuda = torch.cuda.is_available()
    dev = colossalai_b200.accelerator.get_accelerator().get_current_device()
    cfg = get_config(args.model)
    assert not (args.mode == "ring_attn" and not cfg.causal), "ring attention is the causal (decoder) path: --model llama-tiny"
    torch.manual_seed(0)
    model = build_model(cfg)
    optim = HybridAdam(model.parameters(), lr=args.lr)
    if args.mode == "none" or world == 1:
        plugin = HybridParallelPlugin(tp_size=1, pp_size=1, precision="bf16" if cuda else "fp32")
    elif args.mode in ("split_gather", "ring"):                # sequence parallelism inside the tensor-parallel group
        plugin = HybridParallelPlugin(tp_size=world, pp_size=1, precision="bf16" if cuda else "fp32",
                                      enable_sequence_parallelism=True, sequence_parallelism_mode=args.mode)
    else:                                                      # a sequence-parallel group of its own
        plugin = HybridParallelPlugin(tp_size=1, pp_size=1, sp_size=world, precision="bf16" if cuda else "fp32",
                                      enable_sequence_parallelism=True, sequence_parallelism_mode=args.mode)
    booster = Booster(plugin=plugin)
    model, optim, *_ = booster.boost(model, optim)
    gen = torch.Generator().manual_seed(7)                     # every rank of the model replica sees the same batch
    for step in range(args.steps):
        start = torch.randint(0, cfg.vocab_size, (args.batch, 1), generator=gen)
        ids = ((start + torch.arange(args.seq)[None] * 3) % cfg.vocab_size).to(dev)
        out = model(input_ids=ids, labels=ids)
        booster.backward(out["loss"], optim)
        optim.step()
        optim.zero_grad()
        if rank == 0:
            print(f"step {step}: loss {out['loss'].item():.4f} (mode {args.mode}, world {world}, "
                  f"{cfg.model_type} seq {args.seq})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
