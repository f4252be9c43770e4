"""Hybrid-parallel checkpoints: save under one layout (TP2 x PP2), load under others (TP2 x DP2, PP2 x DP2, TP1) — the
HF-style sharded files are layout independent (TP shards gathered, vocab padding removed, PP stage files merged by the
index) and the optimizer states re-shard with the parameters (reference: tests/test_checkpoint_io/
test_hybrid_parallel_plugin_checkpoint_io.py)."""
import os

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.plugin import HybridParallelPlugin
from colossalai_b200.models import build_model, get_config
from colossalai_b200.nn.optimizer import FusedAdam
from colossalai_b200.parallel import comm
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _full_state(model):
    out = {}
    for name, p in model.unwrap().named_parameters():
        if p is None:
            continue
        if hasattr(p, "gather_fn"):
            t = p.gather_fn(p)
        elif hasattr(p, "dist_shard"):
            t = comm.all_gather(p.detach(), p.dist_shard[0], p.dist_shard[1])
        else:
            t = p.detach()
        out[name] = t.clone()
    return out


def _build(plugin_kw, seed, precision="fp32"):
    torch.manual_seed(seed)
    model = build_model(get_config("llama-tiny", vocab_size=500))        # 500 is padded to a multiple of 64 x tp
    opt = FusedAdam(model.parameters(), lr=1e-2)
    plugin = HybridParallelPlugin(precision=precision, num_microbatches=2 if plugin_kw.get("pp_size", 1) > 1 else None,
                                  **plugin_kw)
    booster = Booster(plugin=plugin)
    model, opt, *_ = booster.boost(model, opt)
    return booster, plugin, model, opt


def _step(booster, plugin, model, opt, ids):
    if plugin.pp_size > 1:
        booster.execute_pipeline(iter([{"input_ids": ids, "labels": ids}]), model, lambda o, b: o["loss"], opt,
                                 return_loss=True)
    else:
        booster.backward(model(input_ids=ids, labels=ids)["loss"], opt)
    opt.step()
    opt.zero_grad()


def _worker(rank, world_size, port, tmp):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    ids = torch.randint(0, 500, (2, 16), generator=torch.Generator().manual_seed(3))
    booster, plugin, model, opt = _build(dict(tp_size=2, pp_size=2), seed=1)
    _step(booster, plugin, model, opt, ids)                                # make the optimizer state non-trivial
    ckpt_m, ckpt_o = os.path.join(tmp, "model"), os.path.join(tmp, "optim")
    booster.save_model(model, ckpt_m, shard=True, size_per_shard=1)
    booster.save_optimizer(opt, ckpt_o, shard=True, size_per_shard=1)
    dist.barrier()
    # every pipeline stage only holds its own layers: collect the union over the pp group for the comparison
    src_state = _full_state(model)
    gathered = [None] * world_size
    dist.all_gather_object(gathered, {k: v for k, v in src_state.items()})
    src_all = {}
    for d in gathered:
        src_all.update(d)
    _step(booster, plugin, model, opt, ids)                                # continue one more step under the old layout
    after = [None] * world_size
    dist.all_gather_object(after, _full_state(model))
    cont_all = {}
    for d in after:
        cont_all.update(d)
    for layout in (dict(tp_size=2, pp_size=1), dict(tp_size=1, pp_size=2), dict(tp_size=1, pp_size=1)):
        b2, p2, m2, o2 = _build(layout, seed=99)                           # different init: everything must come from disk
        b2.load_model(m2, ckpt_m)
        b2.load_optimizer(o2, ckpt_o)
        got = [None] * world_size
        dist.all_gather_object(got, _full_state(m2))
        merged = {}
        for d in got:
            merged.update(d)
        assert set(merged) == set(src_all), (layout, set(src_all) ^ set(merged))
        for k, v in src_all.items():
            a = merged[k][: v.shape[0]] if merged[k].shape != v.shape else merged[k]
            b = v[: a.shape[0]] if a.shape != v.shape else v
            torch.testing.assert_close(a, b, msg=lambda m: f"{layout} {k}: {m}")
        # optimizer state came along: one more step lands where the original run landed
        _step(b2, p2, m2, o2, ids)
        got2 = [None] * world_size
        dist.all_gather_object(got2, _full_state(m2))
        merged2 = {}
        for d in got2:
            merged2.update(d)
        for k, v in cont_all.items():
            a = merged2[k][: v.shape[0]] if merged2[k].shape != v.shape else merged2[k]
            b = v[: a.shape[0]] if a.shape != v.shape else v
            torch.testing.assert_close(a, b, atol=2e-4, rtol=2e-3, msg=lambda m: f"resume {layout} {k}: {m}")
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_hybrid_checkpoint_reshards_across_layouts(tmp_path):
    spawn(_worker, 4, tmp=str(tmp_path))


def _zero_worker(rank, world_size, port, tmp):
    """Same idea with ZeRO-1 (bf16 working params, fp32 master + moments in dp-sharded flat buckets)."""
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    ids = torch.randint(0, 500, (2, 16), generator=torch.Generator().manual_seed(3))
    booster, plugin, model, opt = _build(dict(tp_size=2, pp_size=1, zero_stage=1), seed=1, precision="bf16")
    _step(booster, plugin, model, opt, ids)
    ckpt_m, ckpt_o = os.path.join(tmp, "zmodel"), os.path.join(tmp, "zoptim")
    booster.save_model(model, ckpt_m, shard=True, size_per_shard=1)
    booster.save_optimizer(opt, ckpt_o, shard=True, size_per_shard=1)
    dist.barrier()
    src = _full_state(model)
    _step(booster, plugin, model, opt, ids)
    cont = _full_state(model)
    for layout in (dict(tp_size=1, pp_size=1, zero_stage=1), dict(tp_size=2, pp_size=1, zero_stage=2),
                   dict(tp_size=2, pp_size=1)):          # the last one: plain bf16 AMP optimizer (fp32 masters, no ZeRO)
        b2, p2, m2, o2 = _build(layout, seed=99, precision="bf16")
        b2.load_model(m2, ckpt_m)
        b2.load_optimizer(o2, ckpt_o)
        got = _full_state(m2)
        for k, v in src.items():
            a = got[k][: v.shape[0]] if got[k].shape != v.shape else got[k]
            b = v[: a.shape[0]] if a.shape != v.shape else v
            torch.testing.assert_close(a, b, msg=lambda m: f"zero {layout} {k}: {m}")
        _step(b2, p2, m2, o2, ids)
        got2 = _full_state(m2)
        for k, v in cont.items():
            a = got2[k][: v.shape[0]] if got2[k].shape != v.shape else got2[k]
            b = v[: a.shape[0]] if a.shape != v.shape else v
            # bf16 working copies were re-derived from the loaded weights: the masters differ by < 1 bf16 ulp
            bad = ((a.float() - b.float()).abs() > 1.5e-2 + 2e-2 * b.float().abs()).float().mean().item()
            assert bad <= 0.02, f"zero resume {layout} {k}: {bad:.4f} of the elements differ"
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_hybrid_zero_checkpoint_reshards(tmp_path):
    spawn(_zero_worker, 4, tmp=str(tmp_path))
