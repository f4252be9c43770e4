"""HybridParallelPlugin end-to-end on the CPU/gloo tier: boosted tiny Llama (TP=2 [+SP], bf16/fp32) trains and
matches a single-process fp32 oracle step (reference pattern: tests/test_booster/test_plugin/test_3d_plugin.py)."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.plugin import HybridParallelPlugin
from colossalai_b200.models import build_model
from colossalai_b200.nn.optimizer import FusedAdam
from colossalai_b200.parallel import comm
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _gather_param(p):
    if hasattr(p, "gather_fn"):
        return p.gather_fn(p)
    if hasattr(p, "dist_shard"):
        dim, group = p.dist_shard
        return comm.all_gather(p.detach(), dim, group)
    return p.detach()


def _run(cfg):
    torch.manual_seed(42)
    base = build_model("llama-tiny")
    model = copy.deepcopy(base)
    ref_opt = torch.optim.AdamW(base.parameters(), lr=1e-2, weight_decay=0.0)
    opt = FusedAdam(model.parameters(), lr=1e-2, weight_decay=0.0)
    plugin = HybridParallelPlugin(precision="fp32", max_norm=cfg.get("max_norm", 0.0), **cfg["plugin"])
    booster = Booster(plugin=plugin)
    model, opt, *_ = booster.boost(model, opt)
    dp_rank = plugin.pg_mesh.axis_rank("dp")
    torch.manual_seed(100)
    all_ids = torch.randint(0, 512, (2 * plugin.dp_size, 32))
    for _ in range(2):
        out = model(input_ids=all_ids[2 * dp_rank: 2 * dp_rank + 2], labels=all_ids[2 * dp_rank: 2 * dp_rank + 2])
        booster.backward(out["loss"], opt)
        opt.step()
        opt.zero_grad()
        ref = base(input_ids=all_ids, labels=all_ids)
        ref["loss"].backward()
        if cfg.get("max_norm", 0.0) > 0:
            torch.nn.utils.clip_grad_norm_(base.parameters(), cfg["max_norm"])
        ref_opt.step()
        ref_opt.zero_grad()
    ref_params = dict(base.named_parameters())
    n = 0
    for name, p in model.unwrap().named_parameters():
        full = _gather_param(p)
        r = ref_params[name]
        if full.shape != r.shape:
            full = full[: r.shape[0]]
        torch.testing.assert_close(full, r.detach(), atol=2e-4, rtol=2e-3, msg=lambda m: f"{name} {cfg}: {m}")
        n += 1
    assert n > 10
    del plugin


def _run_accum(plugin_kw):
    """Two micro-batches per optimizer step: the second wgrad accumulates inside the GEMM (in-place path)."""
    torch.manual_seed(42)
    base = build_model("llama-tiny")
    model = copy.deepcopy(base)
    ref_opt = torch.optim.AdamW(base.parameters(), lr=1e-2, weight_decay=0.0)
    opt = FusedAdam(model.parameters(), lr=1e-2, weight_decay=0.0)
    plugin = HybridParallelPlugin(precision="fp32", **plugin_kw)
    booster = Booster(plugin=plugin)
    model, opt, *_ = booster.boost(model, opt)
    flagged = [p for p in model.unwrap().parameters() if getattr(p, "_cb200_inplace_wgrad", False)]
    assert flagged, "in-place wgrad accumulation should be enabled without ZeRO"
    torch.manual_seed(100)
    ids = torch.randint(0, 512, (2, 2 * plugin.dp_size, 32))
    dp_rank = plugin.pg_mesh.axis_rank("dp")
    for step in range(2):
        for mb in range(2):
            mine = ids[mb, dp_rank: dp_rank + 1]
            loss = model(input_ids=mine, labels=mine)["loss"] / 2
            if mb == 0:
                with model.no_sync():
                    opt.backward(loss)
            else:
                booster.backward(loss, opt)
            for r in range(plugin.dp_size):
                x = ids[mb, r: r + 1]
                (base(input_ids=x, labels=x)["loss"] / 2 / plugin.dp_size).backward()
        opt.step()
        opt.zero_grad()
        ref_opt.step()
        ref_opt.zero_grad()
    ref_params = dict(base.named_parameters())
    for name, p in model.unwrap().named_parameters():
        full = _gather_param(p)
        r = ref_params[name]
        if full.shape != r.shape:
            full = full[: r.shape[0]]
        torch.testing.assert_close(full, r.detach(), atol=2e-4, rtol=2e-3, msg=lambda m: f"accum {name}: {m}")
    del plugin


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    _run_accum(dict(tp_size=2, pp_size=1, enable_sequence_parallelism=True, sequence_parallelism_mode="split_gather"))
    _run_accum(dict(tp_size=1, pp_size=1))
    _run(dict(plugin=dict(tp_size=2, pp_size=1)))
    _run(dict(plugin=dict(tp_size=2, pp_size=1, enable_sequence_parallelism=True,
                          sequence_parallelism_mode="split_gather"), max_norm=0.5))
    _run(dict(plugin=dict(tp_size=1, pp_size=1), max_norm=0.5))                       # pure DP=2
    _run(dict(plugin=dict(tp_size=1, pp_size=1, sp_size=2, enable_sequence_parallelism=True,
                          sequence_parallelism_mode="all_to_all")))
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_hybrid_plugin_cpu():
    spawn(_worker, 2)


def _bf16_worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    torch.manual_seed(0)
    model = build_model("llama-tiny")
    opt = FusedAdam(model.parameters(), lr=1e-3)
    booster = Booster(plugin=HybridParallelPlugin(tp_size=2, pp_size=1, precision="bf16", max_norm=1.0))
    model, opt, *_ = booster.boost(model, opt)
    ids = torch.randint(0, 512, (2, 32))
    losses = []
    for _ in range(8):
        out = model(input_ids=ids, labels=ids)
        booster.backward(out["loss"], opt)
        opt.step()
        opt.zero_grad()
        losses.append(out["loss"].item())
    assert losses[-1] < losses[0], losses
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_hybrid_plugin_bf16_learns():
    spawn(_bf16_worker, 2)


if __name__ == "__main__":
    test_hybrid_plugin_cpu()
    test_hybrid_plugin_bf16_learns()
