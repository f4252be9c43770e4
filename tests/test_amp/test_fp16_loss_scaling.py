"""fp16 mixed precision with dynamic loss scaling through every plugin that owns a scaler: an overflowing step is
skipped (parameters untouched, scale backed off), a healthy one follows the fp32 oracle (reference:
tests/test_zero/test_low_level/test_zero1_2.py fp16 branch, amp/naive_amp/mixed_precision_mixin/fp16.py)."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.plugin import GeminiPlugin, HybridParallelPlugin, LowLevelZeroPlugin
from colossalai_b200.models import build_model
from colossalai_b200.nn.optimizer import HybridAdam
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _scale(opt) -> float:
    cur = opt
    for _ in range(4):
        for name in ("mixed_precision_mixin", "mix_precision_mixin", "mixed_precision"):
            m = getattr(cur, name, None)
            if m is not None and hasattr(m, "loss_scale"):
                return float(m.loss_scale)
        cur = getattr(cur, "optim", None)
        if cur is None:
            break
    raise AssertionError("no loss scaler found")


def _weights(model):
    from colossalai_b200.tensor.d_tensor import to_global
    from colossalai_b200.zero.gemini import GeminiDDP

    if isinstance(model, GeminiDDP):
        return {k: v.float() for k, v in model.state_dict(only_rank_0=False, dtype=torch.float32).items()}
    inner = model.unwrap() if hasattr(model, "unwrap") else model
    return {n: to_global(p).detach().float().clone() for n, p in inner.named_parameters()}


def _boost(make_plugin, scale):
    torch.manual_seed(3)
    base = build_model("llama-tiny").float()
    model = copy.deepcopy(base)
    opt = HybridAdam(model.parameters(), lr=2e-3)
    model, opt, *_ = (booster := Booster(plugin=make_plugin(scale))).boost(model, opt)
    return base, model, opt, booster


def _check(tag, make_plugin):
    ids = torch.randint(0, 512, (4, 16), generator=torch.Generator().manual_seed(10))
    # ---- overflow: 2**40 * grad does not fit fp16 -> every step is skipped; with hysteresis 2 the scale backs off from
    # the second overflow on
    base, model, opt, booster = _boost(make_plugin, 2.0 ** 40)
    before = _weights(model)
    for _ in range(3):
        booster.backward(model(input_ids=ids, labels=ids)["loss"], opt)
        opt.step()
        opt.zero_grad()
    after = _weights(model)
    for k, v in before.items():
        assert torch.equal(v, after[k]), f"{tag}: {k} moved on an overflowing step"
    assert _scale(opt) == 2.0 ** 38, f"{tag}: scale {_scale(opt)}"
    # ---- healthy scale: follows the fp32 AdamW oracle
    base, model, opt, booster = _boost(make_plugin, 2.0 ** 10)
    ref_opt = torch.optim.AdamW(base.parameters(), lr=2e-3, weight_decay=0.0)
    for step in range(3):
        loss = model(input_ids=ids, labels=ids)["loss"]
        booster.backward(loss, opt)
        opt.step()
        opt.zero_grad()
        ref = base(input_ids=ids, labels=ids)["loss"]
        ref.backward()
        ref_opt.step()
        ref_opt.zero_grad()
        assert abs(loss.item() - ref.item()) < 3e-2, f"{tag} step {step}: {loss.item()} vs {ref.item()}"
    assert _scale(opt) == 2.0 ** 10
    got = _weights(model)
    for k, v in base.state_dict().items():
        if k in got:
            d = (got[k] - v.float()).abs()
            assert d.mean().item() < 1.5e-3, f"{tag} {k}: mean {d.mean()}"


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    cases = [
        ("zero1", lambda s: LowLevelZeroPlugin(stage=1, precision="fp16", initial_scale=s)),
        ("zero2", lambda s: LowLevelZeroPlugin(stage=2, precision="fp16", initial_scale=s)),
        ("hybrid tp2", lambda s: HybridParallelPlugin(tp_size=2, pp_size=1, precision="fp16", initial_scale=s)),
        ("hybrid zero1", lambda s: HybridParallelPlugin(tp_size=1, pp_size=1, zero_stage=1, precision="fp16",
                                                        initial_scale=s)),
        ("gemini", lambda s: GeminiPlugin(precision="fp16", initial_scale=s, min_chunk_size_m=0.01, search_range_m=1)),
    ]
    for tag, fn in cases:
        _check(tag, fn)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_fp16_dynamic_loss_scaling_all_plugins():
    spawn(_worker, 2)
