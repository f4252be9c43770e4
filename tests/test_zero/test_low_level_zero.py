"""ZeRO-1/2 vs a DDP-equivalent single-process oracle (reference: tests/test_zero/test_low_level/test_zero1_2.py,
test_grad_acc.py, test_zero_ckpt.py) on gloo, dp=2."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.plugin import LowLevelZeroPlugin
from colossalai_b200.models import build_model
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _run(stage, accum):
    torch.manual_seed(3)
    base = build_model("llama-tiny")
    model = copy.deepcopy(base)
    ref_opt = torch.optim.AdamW(base.parameters(), lr=1e-2, weight_decay=0.01)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.01)
    plugin = LowLevelZeroPlugin(stage=stage, precision="fp32", max_norm=0.7, reduce_bucket_size_in_m=1,
                                overlap_communication=False)
    booster = Booster(plugin=plugin)
    model, opt, *_ = booster.boost(model, opt)
    rank = dist.get_rank()
    torch.manual_seed(5)
    for _ in range(2):
        batches = [torch.randint(0, 512, (4, 16)) for _ in range(accum)]
        for i, ids in enumerate(batches):
            mine = ids[2 * rank: 2 * rank + 2]
            loss = model(input_ids=mine, labels=mine)["loss"] / accum
            if i < accum - 1:
                with booster.no_sync(model, opt):
                    booster.backward(loss, opt)
            else:
                booster.backward(loss, opt)
            (base(input_ids=ids, labels=ids)["loss"] / accum).backward()
        opt.step()
        opt.zero_grad()
        torch.nn.utils.clip_grad_norm_(base.parameters(), 0.7)
        ref_opt.step()
        ref_opt.zero_grad()
    for (n, p), (_, r) in zip(model.unwrap().named_parameters(), base.named_parameters()):
        torch.testing.assert_close(p.detach(), r.detach(), atol=2e-4, rtol=2e-3, msg=lambda m: f"z{stage} {n}: {m}")
    # optimizer state round trip
    sd = opt.state_dict()
    assert len(sd["state"]) == len(list(base.parameters()))
    opt.load_state_dict(sd)


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    _run(1, 1)
    _run(1, 2)
    _run(2, 1)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_low_level_zero_cpu():
    spawn(_worker, 2)


if __name__ == "__main__":
    test_low_level_zero_cpu()


def _untouched_worker(rank, world_size, port):
    """`skip_untouched_params=True`: a parameter that receives no gradient on any rank in a step keeps its value AND its
    momentum, exactly like `torch.optim` (SGD with momentum: no step counters, so the match is exact); the default
    steps it with a zero gradient (momentum keeps moving it)."""
    import copy

    import torch.nn as nn

    import colossalai_b200
    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import LowLevelZeroPlugin

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(16, 32), nn.Linear(32, 16)
            self.sometimes, self.never = nn.Linear(16, 16), nn.Linear(16, 16)

        def forward(self, x, use):
            h = self.b(torch.relu(self.a(x)))
            return h + self.sometimes(x) if use else h

    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    for stage in (1, 2):
        diffs = {}
        for skip in (True, False):
            torch.manual_seed(0)
            base = Net()
            model = copy.deepcopy(base)
            opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
            ref_opt = torch.optim.SGD(base.parameters(), lr=0.1, momentum=0.9)
            booster = Booster(plugin=LowLevelZeroPlugin(stage=stage, precision="fp32", skip_untouched_params=skip))
            model, opt, *_ = booster.boost(model, opt)
            g = torch.Generator().manual_seed(5)
            for step in range(6):
                xs = [torch.randn(4, 16, generator=g) for _ in range(world_size)]
                use = step % 3 == 0
                booster.backward(model(xs[rank], use).square().mean(), opt)
                opt.step()
                opt.zero_grad()
                ref_opt.zero_grad()
                for x in xs:
                    (base(x, use).square().mean() / world_size).backward()
                ref_opt.step()
            diffs[skip] = {n: float((p.detach() - q.detach()).abs().max())
                           for (n, p), (_, q) in zip(model.unwrap().named_parameters(), base.named_parameters())}
        assert max(diffs[True].values()) < 1e-5, (stage, diffs[True])
        assert diffs[False]["sometimes.weight"] > 1e-3 and diffs[False]["never.weight"] == 0.0, (stage, diffs[False])
    dist.destroy_process_group()


@pytest.mark.dist
def test_zero_skip_untouched_params_matches_torch_semantics():
    spawn(_untouched_worker, 2)
