"""Expert-parallel MoE training with MoeHybridParallelPlugin (reference: examples/language/mixtral/train.py).

    torchrun --nproc-per-node 8 examples/language/mixtral/train.py -c mixtral-tiny --ep 8 --zero 1
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))

import colossalai_b200  # noqa: E402
from colossalai_b200.booster import Booster  # noqa: E402
from colossalai_b200.booster.plugin import MoeHybridParallelPlugin  # noqa: E402
from colossalai_b200.models import build_model, get_config  # noqa: E402
from colossalai_b200.nn.optimizer import FusedAdam  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", default="mixtral-tiny")
    ap.add_argument("--ep", type=int, default=2)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--zero", type=int, default=1)
    ap.add_argument("-b", "--batch_size", type=int, default=2)
    ap.add_argument("-l", "--max_length", type=int, default=128)
    ap.add_argument("-s", "--num_steps", type=int, default=5)
    ap.add_argument("--save_dir", default=None)
    args = ap.parse_args()
    colossalai_b200.launch_from_torch()
    cfg = get_config(args.config)
    plugin = MoeHybridParallelPlugin(ep_size=args.ep, tp_size=args.tp, pp_size=args.pp, zero_stage=args.zero,
                                     precision="bf16", max_norm=1.0,
                                     microbatch_size=1 if args.pp > 1 else None)
    booster = Booster(plugin=plugin)
    model = build_model(cfg)
    optimizer = FusedAdam(model.parameters(), lr=1e-4, weight_decay=0.01)
    model, optimizer, *_ = booster.boost(model, optimizer)
    dev = colossalai_b200.accelerator.get_accelerator().get_current_device()
    g = torch.Generator().manual_seed(1234 + plugin.pg_mesh.axis_rank("dp"))
    for step in range(args.num_steps):
        ids = torch.randint(0, cfg.vocab_size, (args.batch_size, args.max_length), generator=g).to(dev)
        out = model(input_ids=ids, labels=ids)
        loss = out["loss"]
        aux = [m.aux_loss for m in model.unwrap().modules() if getattr(m, "aux_loss", None) is not None]
        if aux:
            loss = loss + sum(aux)
        booster.backward(loss, optimizer)
        optimizer.step()
        optimizer.zero_grad()
        if dist.get_rank() == 0:
            print(f"step {step}: loss {loss.item():.4f}")
    if args.save_dir:
        booster.save_model(model, args.save_dir, shard=True)
    dist.barrier()
    colossalai_b200.initialize.shutdown()


if __name__ == "__main__":
    main()
