"""GRPO / DAPO on a synthetic verifiable task, producer + consumer on the ranks of one torchrun job.

    python applications/ColossalChat/examples/train_grpo.py --steps 20                 # single process
    torchrun --nproc-per-node 2 applications/ColossalChat/examples/train_grpo.py       # rank 0 produces, rank 1 trains
    torchrun --nproc-per-node 3 applications/ColossalChat/examples/train_grpo.py --producers 2 \
        --eval_interval 5 --save_dir /tmp/grpo --save_interval 10 --rollout_log /tmp/grpo/rollouts.jsonl
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
    ap.add_argument("--producers", type=int, default=1, help="producer ranks (multi-process): ranks 0..P-1 generate")
    ap.add_argument("--eval_interval", type=int, default=0)
    ap.add_argument("--save_dir", default=None)
    ap.add_argument("--save_interval", type=int, default=0)
    ap.add_argument("--rollout_log", default=None)
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
    n_prod = args.producers if multi else 1
    if not multi or rank < n_prod:
        sampler = build_model(cfg).float().to(dev)
        producer = Producer(ModelRolloutBackend(sampler, dict(max_new_tokens=args.max_new_tokens)), prompts,
                            args.num_generations, producer_idx=rank if multi else 0, num_producers=n_prod,
                            rollout_log=args.rollout_log if rank == 0 else None)
    if not multi or rank >= n_prod:
        policy = build_model(cfg).float().to(dev)
        consumer = GRPOConsumer(policy, torch.optim.AdamW(policy.parameters(), lr=args.lr), reward,
                                num_generations=args.num_generations, clip_eps_high=0.28, loss_variation="token_level")
    hist = launch_distributed(producer, consumer, args.steps, sync_every=2, producer_ranks=tuple(range(n_prod)),
                              eval_dataloaders={"held_out": prompts} if args.eval_interval else None,
                              eval_interval=args.eval_interval, eval_reward_fn=reward, save_dir=args.save_dir,
                              save_interval=args.save_interval)
    if producer is not None and rank == 0 and getattr(producer, "last_eval", None):
        print("last evaluation:", producer.last_eval)
    if consumer is not None and (not multi or rank == n_prod):
        for i, h in enumerate(hist):
            print(f"step {i:3d} reward {h['reward']:.3f} loss {h.get('loss', 0.0):+.4f} kept {h['kept']:.2f}")


if __name__ == "__main__":
    main()
