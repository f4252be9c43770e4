"""Graph-level auto parallelism on a transformer block (reference: examples/tutorial/auto_parallel/
auto_parallel_with_resnet.py).

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 4 auto_parallel_block.py --mesh 2 2

Traces the block, lets the ILP pick a strategy per node for the given logical mesh, applies it and trains a few steps;
rank 0 prints the chosen plan.
"""
import argparse
import math

import torch
import torch.distributed as dist
import torch.nn as nn

import colossalai_b200
from colossalai_b200.auto_parallel.tensor_shard import SolverOptions, initialize_device_mesh, initialize_model


class Block(nn.Module):
    def __init__(self, h: int, f: int, nh: int) -> None:
        super().__init__()
        self.ln1, self.ln2 = nn.LayerNorm(h), nn.LayerNorm(h)
        self.q, self.k, self.v, self.o = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
        self.up, self.down, self.act = nn.Linear(h, f), nn.Linear(f, h), nn.GELU()
        self.nh, self.hd = nh, h // nh

    def forward(self, x):
        B, S, H = x.shape[0], x.shape[1], x.shape[2]
        y = self.ln1(x)
        q = self.q(y).view(B, S, self.nh, self.hd).transpose(1, 2)
        k = self.k(y).view(B, S, self.nh, self.hd).transpose(1, 2)
        v = self.v(y).view(B, S, self.nh, self.hd).transpose(1, 2)
        p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.hd), dim=-1)
        x = x + self.o(torch.matmul(p, v).transpose(1, 2).reshape(B, S, H))
        return x + self.down(self.act(self.up(self.ln2(x))))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, nargs="+", default=None, help="logical mesh shape, e.g. 2 4")
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--memory-budget-gb", type=float, default=-1.0)
    args = ap.parse_args()
    colossalai_b200.launch_from_torch()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    torch.manual_seed(0)
    model = Block(args.hidden, 4 * args.hidden, max(args.hidden // 128, 1)).to(dev)
    mesh = initialize_device_mesh(logical_mesh_shape=tuple(args.mesh) if args.mesh else None)
    meta = {"x": torch.empty(args.batch, args.seq, args.hidden, device="meta")}
    budget = args.memory_budget_gb * 2**30 if args.memory_budget_gb > 0 else -1.0
    gm, plan, specs = initialize_model(model, meta, mesh, memory_budget=budget, solver_options=SolverOptions(),
                                       return_solution=True)
    if dist.get_rank() == 0:
        print("plan:", {k: v for k, v in plan.items() if "@" in v or "S" in v.split("[")[-1]} or "everything replicated")
        print("sharded parameters:", specs)
    opt = torch.optim.AdamW(gm.parameters(), lr=1e-3)
    x = torch.randn(args.batch, args.seq, args.hidden, device=dev)
    for step in range(3):
        loss = gm(x).float().pow(2).mean()
        loss.backward()
        opt.step()
        opt.zero_grad()
        if dist.get_rank() == 0:
            print(f"step {step}: loss {loss.item():.5f}")


if __name__ == "__main__":
    main()
