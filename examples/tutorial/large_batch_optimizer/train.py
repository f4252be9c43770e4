"""Large-batch training with layer-wise adaptive optimizers (LARS / LAMB) under data parallelism.

    python examples/tutorial/large_batch_optimizer/train.py --optimizer lars --batch 512
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/tutorial/large_batch_optimizer/train.py --optimizer lamb \
        --plugin zero1

A small residual CNN on a synthetic 10-class image task (every class tints the noise image with its own colour, so the
problem is learnable); the learning rate follows the linear-scaling rule with warm-up, which is where plain SGD gets
unstable at large batch sizes and the per-layer trust ratio of LARS / LAMB does not.
Parity: reference `examples/tutorial/large_batch_optimizer/train.py` (ResNet on CIFAR-10 with `Lars` / `Lamb`).
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))

import colossalai_b200  # noqa: E402
from colossalai_b200.booster import Booster  # noqa: E402
from colossalai_b200.booster.plugin import LowLevelZeroPlugin, TorchDDPPlugin  # noqa: E402
from colossalai_b200.nn.lr_scheduler import CosineAnnealingWarmupLR  # noqa: E402
from colossalai_b200.nn.optimizer import Lamb, Lars  # noqa: E402
from colossalai_b200.testing import free_port  # noqa: E402


class Block(nn.Module):
    def __init__(self, c: int) -> None:
        super().__init__()
        self.a, self.b = nn.Conv2d(c, c, 3, padding=1, bias=False), nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.na, self.nb = nn.GroupNorm(4, c), nn.GroupNorm(4, c)

    def forward(self, x):
        return F.relu(x + self.nb(self.b(F.relu(self.na(self.a(x))))))


class SmallResNet(nn.Module):
    def __init__(self, width: int = 32, classes: int = 10) -> None:
        super().__init__()
        self.stem = nn.Conv2d(3, width, 3, padding=1)
        self.blocks = nn.Sequential(Block(width), nn.AvgPool2d(2), Block(width), nn.AvgPool2d(2), Block(width))
        self.head = nn.Linear(width, classes)

    def forward(self, x):
        return self.head(self.blocks(F.relu(self.stem(x))).mean((2, 3)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--optimizer", default="lars", choices=["lars", "lamb", "sgd"])
    ap.add_argument("--plugin", default="ddp", choices=["ddp", "zero1"])
    ap.add_argument("--batch", type=int, default=256, help="GLOBAL batch size")
    ap.add_argument("--base_lr", type=float, default=None, help="learning rate at batch 256 (linear scaling rule)")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--image", type=int, default=16)
    args = ap.parse_args()
    if "RANK" in os.environ:
        colossalai_b200.launch_from_torch(backend="nccl" if torch.cuda.is_available() else "gloo")
    else:
        colossalai_b200.launch(0, 1, "127.0.0.1", free_port(), verbose=False)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = colossalai_b200.accelerator.get_accelerator().get_current_device()
    torch.manual_seed(0)
    model = SmallResNet().to(dev)
    base = args.base_lr if args.base_lr is not None else {"lars": 2.0, "lamb": 1e-2, "sgd": 0.05}[args.optimizer]
    lr = base * args.batch / 256
    if args.optimizer == "lars":
        optim = Lars(model.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
    elif args.optimizer == "lamb":
        optim = Lamb(model.parameters(), lr=lr, weight_decay=1e-2)
    else:
        optim = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
    sched = CosineAnnealingWarmupLR(optim, total_steps=args.steps, warmup_steps=args.warmup)
    plugin = TorchDDPPlugin() if args.plugin == "ddp" else LowLevelZeroPlugin(stage=1, precision="fp32")
    booster = Booster(plugin=plugin)
    model, optim, _, _, sched = booster.boost(model, optim, lr_scheduler=sched)
    tints = torch.randn(10, 3, generator=torch.Generator().manual_seed(1))
    gen = torch.Generator().manual_seed(100 + rank)
    local = max(args.batch // world, 1)
    for step in range(args.steps):
        y = torch.randint(0, 10, (local,), generator=gen)
        x = (torch.randn(local, 3, args.image, args.image, generator=gen) + 0.5 * tints[y][:, :, None, None]).to(dev)
        y = y.to(dev)
        logits = model(x)
        loss = F.cross_entropy(logits, y)
        booster.backward(loss, optim)
        optim.step()
        optim.zero_grad()
        sched.step()
        stats = torch.stack([loss.detach(), (logits.argmax(-1) == y).float().mean()])
        dist.all_reduce(stats)
        stats /= world
        if rank == 0 and (step % 10 == 0 or step == args.steps - 1):
            print(f"step {step:3d} loss {stats[0].item():.4f} acc {stats[1].item():.3f} "
                  f"lr {optim.param_groups[0]['lr']:.4g} ({args.optimizer}, global batch {local * world})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
