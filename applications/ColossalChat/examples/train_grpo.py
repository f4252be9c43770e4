"""GRPO / DAPO on a synthetic verifiable task, producer + consumer on the ranks of one torchrun job.

    python applications/ColossalChat/examples/train_grpo.py --steps 20                 # single process
    torchrun --nproc-per-node 2 applications/ColossalChat/examples/train_grpo.py       # rank 0 produces, rank 1 trains
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))

from coati.dataset import DataCollatorForPromptDataset, ListDataset  # noqa: E402
from coati.distributed import GRPOConsumer, ModelRolloutBackend, Producer, launch_distributed  # noqa: E402

import colossalai_b200  # noqa: E402
from colossalai_b200.models import build_model, get_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-tiny")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--num_generations", type=int, default=8)
    ap.add_argument("--max_new_tokens", type=int, default=8)
    ap.add_argument("--lr", type=float, default=3e-3)
    ap.add_argument("--target_token", type=int, default=7)
    args = ap.parse_args()
    multi = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    if multi:
        colossalai_b200.launch_from_torch(backend="nccl" if torch.cuda.is_available() else "gloo")
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    rank = dist.get_rank() if multi else 0
    torch.manual_seed(0)
    cfg = get_config(args.model, vocab_size=64)
    prompts = torch.utils.data.DataLoader(ListDataset([{"input_ids": [1, 10 + i, 20 + i]} for i in range(8)]),
                                          batch_size=4, collate_fn=DataCollatorForPromptDataset())

    def reward(seq, prompt_len, **_):
        return (seq[:, prompt_len:] == args.target_token).float().mean(-1)

    producer = consumer = None
    if not multi or rank == 0:
        sampler = build_model(cfg).float().to(dev)
        producer = Producer(ModelRolloutBackend(sampler, dict(max_new_tokens=args.max_new_tokens)), prompts,
                            args.num_generations)
    if not multi or rank != 0:
        policy = build_model(cfg).float().to(dev)
        consumer = GRPOConsumer(policy, torch.optim.AdamW(policy.parameters(), lr=args.lr), reward,
                                num_generations=args.num_generations, clip_eps_high=0.28, loss_variation="token_level")
    hist = launch_distributed(producer, consumer, args.steps, sync_every=2)
    if consumer is not None and (not multi or rank == 1):
        for i, h in enumerate(hist):
            print(f"step {i:3d} reward {h['reward']:.3f} loss {h.get('loss', 0.0):+.4f} kept {h['kept']:.2f}")


if __name__ == "__main__":
    main()
