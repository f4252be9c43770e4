"""ZeRO-1/2 (LowLevelZeroPlugin) and Gemini checkpoints: save model + optimizer mid-run (sharded and unsharded), load
into a freshly built run, and the next step lands exactly where the uninterrupted run lands (reference:
tests/test_checkpoint_io/test_low_level_zero_checkpoint_io.py, test_gemini_checkpoint_io.py)."""
import os

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.plugin import GeminiPlugin, LowLevelZeroPlugin
from colossalai_b200.models import build_model
from colossalai_b200.nn.optimizer import HybridAdam
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _make(plugin_fn, seed):
    torch.manual_seed(seed)
    model = build_model("llama-tiny")
    opt = HybridAdam(model.parameters(), lr=1e-2)
    booster = Booster(plugin=plugin_fn())
    model, opt, *_ = booster.boost(model, opt)
    return booster, model, opt


def _step(booster, model, opt, ids):
    booster.backward(model(input_ids=ids, labels=ids)["loss"], opt)
    opt.step()
    opt.zero_grad()


def _logits(model, ids):
    model.eval()
    with torch.no_grad():
        out = model(input_ids=ids)["logits"].float().clone()
    model.train()
    return out


def _roundtrip(plugin_fn, tmp, tag, shard):
    rank = dist.get_rank()
    ids = torch.randint(0, 512, (2, 16), generator=torch.Generator().manual_seed(5 + rank))
    probe = torch.randint(0, 512, (1, 16), generator=torch.Generator().manual_seed(77))
    booster, model, opt = _make(plugin_fn, seed=1)
    _step(booster, model, opt, ids)
    m_path = os.path.join(tmp, f"{tag}_model" + ("" if shard else ".pt"))
    o_path = os.path.join(tmp, f"{tag}_optim" + ("" if shard else ".pt"))
    booster.save_model(model, m_path, shard=shard, size_per_shard=1)
    booster.save_optimizer(opt, o_path, shard=shard, size_per_shard=1)
    dist.barrier()
    saved_logits = _logits(model, probe)
    _step(booster, model, opt, ids)
    cont_logits = _logits(model, probe)
    b2, m2, o2 = _make(plugin_fn, seed=123)           # different init
    b2.load_model(m2, m_path)
    b2.load_optimizer(o2, o_path)
    torch.testing.assert_close(_logits(m2, probe), saved_logits, atol=2e-2, rtol=2e-2, msg=lambda m: f"{tag} load: {m}")
    _step(b2, m2, o2, ids)
    torch.testing.assert_close(_logits(m2, probe), cont_logits, atol=3e-2, rtol=3e-2, msg=lambda m: f"{tag} resume: {m}")


def _worker(rank, world_size, port, tmp):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    for stage, shard in ((1, True), (2, False)):
        _roundtrip(lambda: LowLevelZeroPlugin(stage=stage, precision="bf16"), tmp, f"zero{stage}_{int(shard)}", shard)
    for shard in (True, False):
        _roundtrip(lambda: GeminiPlugin(precision="bf16", placement_policy="static", min_chunk_size_m=0.01, search_range_m=1), tmp, f"gemini_{int(shard)}", shard)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_zero_and_gemini_checkpoint_resume(tmp_path):
    spawn(_worker, 2, tmp=str(tmp_path))
