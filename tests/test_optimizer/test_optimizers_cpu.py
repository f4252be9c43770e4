"""Optimizers against torch references on CPU (reference: tests/test_optimizer/test_adam_kernel.py, test_nvme.py,
test_lr_scheduler.py)."""
import math
import os

import pytest
import torch
import torch.nn as nn

from colossalai_b200.nn.lr_scheduler import CosineAnnealingWarmupLR, LinearWarmupLR
from colossalai_b200.nn.optimizer import CPUAdam, FusedAdam, HybridAdam


def _pair(seed=0):
    torch.manual_seed(seed)
    a = nn.Sequential(nn.Linear(16, 32), nn.Tanh(), nn.Linear(32, 4))
    b = nn.Sequential(nn.Linear(16, 32), nn.Tanh(), nn.Linear(32, 4))
    b.load_state_dict(a.state_dict())
    return a, b


@pytest.mark.parametrize("cls", [CPUAdam, HybridAdam, FusedAdam])
@pytest.mark.parametrize("adamw", [True, False])
def test_adam_family_matches_torch(cls, adamw):
    a, b = _pair()
    kw = dict(lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    try:
        opt = cls(a.parameters(), adamw_mode=adamw, **kw)
    except TypeError:
        opt = cls(a.parameters(), **kw)
        if not adamw:
            pytest.skip("optimizer has no adamw_mode switch")
    ref = (torch.optim.AdamW if adamw else torch.optim.Adam)(b.parameters(), **kw)
    x = torch.randn(8, 16)
    for _ in range(5):
        for m, o in ((a, opt), (b, ref)):
            o.zero_grad()
            m(x).square().mean().backward()
            o.step()
    for p, q in zip(a.parameters(), b.parameters()):
        torch.testing.assert_close(p, q, rtol=1e-4, atol=1e-6)


def test_nvme_offload_adam(tmp_path):
    a, b = _pair(1)
    opt = CPUAdam(a.parameters(), lr=1e-2, nvme_offload_fraction=1.0, nvme_offload_dir=str(tmp_path))
    ref = torch.optim.AdamW(b.parameters(), lr=1e-2, weight_decay=0.0)
    x = torch.randn(8, 16)
    for _ in range(3):
        for m, o in ((a, opt), (b, ref)):
            o.zero_grad()
            m(x).square().mean().backward()
            o.step()
    for p, q in zip(a.parameters(), b.parameters()):
        torch.testing.assert_close(p, q, rtol=1e-4, atol=1e-6)
    assert any(os.scandir(tmp_path)), "optimizer states should have been written to the offload directory"


def test_lr_schedulers():
    p = [nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD(p, lr=1.0)
    s = CosineAnnealingWarmupLR(opt, total_steps=20, warmup_steps=5, eta_min=0.1)
    lrs = []
    for _ in range(20):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        s.step()
    assert lrs[0] < lrs[4] <= 1.0 + 1e-9 and lrs[5] >= lrs[10] >= lrs[19] >= 0.1 - 1e-6
    opt = torch.optim.SGD(p, lr=1.0)
    s = LinearWarmupLR(opt, total_steps=10, warmup_steps=4)
    vals = []
    for _ in range(10):
        vals.append(opt.param_groups[0]["lr"])
        opt.step()
        s.step()
    assert vals[1] > vals[0] and vals[-1] < vals[4]
