"""Regression oracles for gradient syncs that must happen BEFORE the ZeRO reduction (CPU / gloo tier):

* ZeRO-1/2 + Megatron sequence parallelism (`split_gather`): norm weights only see a sequence slice, their partial
  gradients must be summed over the tp group although the ZeRO hook consumes `p.grad` immediately
  (reference: `colossalai/shardformer/layer/utils.py:75-127` runs before `HybridParallelZeroOptimizer` reduces).
* PP + ZeRO-1 + tied embeddings: the tied weight's gradient is the sum of the first and last stage's contributions
  (reference: `colossalai/booster/plugin/hybrid_parallel_plugin.py:88-137`).
* expert-parallel 1/ep gradient scaling under gradient accumulation without ZeRO (each micro-batch scaled once).
"""
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


def _reduced_grads(model, opt):
    """name -> synchronised gradient, read from `p.grad` (no ZeRO) or from the reduced flat bucket shards (ZeRO with a
    data-parallel size of 1: the shard is the whole bucket)."""
    names = {id(p): n for n, p in model.unwrap().named_parameters() if p is not None}
    out = {}
    buckets = getattr(opt, "buckets", None)
    if buckets is None:
        for n, p in model.unwrap().named_parameters():
            if p is not None and p.grad is not None:
                out[n] = p.grad.detach().float().clone()
        return out
    for b in buckets:
        assert b.ws == 1 and b.grad_shard is not None
        for p, o in zip(b.params, b.offsets):
            out[names[id(p)]] = b.grad_shard[o:o + p.numel()].view(p.shape).detach().float().clone()
    return out


def _zero_sp():
    """bf16, tp=2 + split_gather: the synchronised gradients must not depend on the ZeRO stage / overlap setting."""
    results = {}
    for tag, kw in [("zero0", dict(zero_stage=0)), ("zero1", dict(zero_stage=1, overlap_communication=False)),
                    ("zero1+overlap", dict(zero_stage=1, overlap_communication=True)),
                    ("zero2", dict(zero_stage=2, overlap_communication=False))]:
        torch.manual_seed(42)
        model = build_model("llama-tiny")
        opt = FusedAdam(model.parameters(), lr=1e-2, weight_decay=0.0)
        plugin = HybridParallelPlugin(tp_size=2, pp_size=1, precision="bf16", enable_sequence_parallelism=True,
                                      sequence_parallelism_mode="split_gather", **kw)
        if tag == "zero1+overlap":
            plugin.zero_config["overlap_communication"] = True
        booster = Booster(plugin=plugin)
        model, opt, *_ = booster.boost(model, opt)
        if tag == "zero1+overlap":          # the CPU tier has no side stream; force the early-reduce code path
            opt._overlap_communication = True
        torch.manual_seed(100)
        ids = torch.randint(0, 512, (2, 32))
        booster.backward(model(input_ids=ids, labels=ids)["loss"], opt)
        results[tag] = _reduced_grads(model, opt)
        opt.zero_grad()
        del plugin
    ref = results["zero0"]
    n_norm = 0
    for tag, got in results.items():
        if tag == "zero0":
            continue
        for name, g in got.items():
            scale = ref[name].abs().max().item() + 1e-8
            err = (g - ref[name]).abs().max().item() / scale
            assert err < 2e-2, f"{tag} {name}: relative gradient error {err:.3f} vs the non-ZeRO run"
            n_norm += "norm" in name
    assert n_norm >= 9


def _pp_zero_tied():
    """bf16, pp=2, gpt2-tiny (tied embedding / head): ZeRO-1 must see the SUM of both stages' tied gradients."""
    results = {}
    for zero_stage in (0, 1):
        torch.manual_seed(7)
        model = build_model("gpt2-tiny")
        opt = FusedAdam(model.parameters(), lr=1e-2, weight_decay=0.0)
        plugin = HybridParallelPlugin(tp_size=1, pp_size=2, precision="bf16", num_microbatches=4,
                                      zero_stage=zero_stage)
        booster = Booster(plugin=plugin)
        model, opt, *_ = booster.boost(model, opt)
        torch.manual_seed(11)
        ids = torch.randint(0, 512, (4, 16))
        booster.execute_pipeline(iter([{"input_ids": ids, "labels": ids}]), model, lambda o, b: o["loss"], opt,
                                 return_loss=True)
        grads = _reduced_grads(model, opt)
        tied_name = "model.embed_tokens.weight" if dist.get_rank() == 0 else "lm_head.weight"
        assert tied_name in grads, list(grads)
        results[zero_stage] = grads[tied_name]
        for _ in range(2):
            opt.step()
            opt.zero_grad()
            booster.execute_pipeline(iter([{"input_ids": ids, "labels": ids}]), model, lambda o, b: o["loss"], opt,
                                     return_loss=True)
        opt.step()
        tied = dict(model.unwrap().named_parameters())[tied_name].detach().float().clone()
        both = [torch.empty_like(tied) for _ in range(2)]
        dist.all_gather(both, tied)
        # the two copies of the tied weight never drift apart
        torch.testing.assert_close(both[0], both[1], atol=0, rtol=0, msg=lambda m: f"zero{zero_stage}: {m}")
        g = results[zero_stage]
        gb = [torch.empty_like(g) for _ in range(2)]
        dist.all_gather(gb, g)
        torch.testing.assert_close(gb[0], gb[1], atol=0, rtol=0, msg=lambda m: f"zero{zero_stage} tied grad: {m}")
        del plugin
    scale = results[0].abs().max().item()
    assert (results[1] - results[0]).abs().max().item() < 2e-2 * scale


def _moe_accum():
    from colossalai_b200.booster.plugin import MoeHybridParallelPlugin
    from colossalai_b200.tensor.moe_tensor import is_moe_tensor

    grads = {}
    for n_mb in (1, 2):
        torch.manual_seed(3)
        model = build_model("mixtral-tiny")
        opt = torch.optim.SGD(model.parameters(), lr=0.0)
        plugin = MoeHybridParallelPlugin(ep_size=2, tp_size=1, pp_size=1, zero_stage=0, precision="fp32")
        booster = Booster(plugin=plugin)
        model, opt, *_ = booster.boost(model, opt)
        torch.manual_seed(50 + dist.get_rank())
        ids = torch.randint(0, 512, (1, 16))
        for mb in range(n_mb):
            loss = model(input_ids=ids, labels=ids)["loss"]
            if mb < n_mb - 1:
                with booster.no_sync(model, opt):
                    opt.backward(loss)
            else:
                booster.backward(loss, opt)
        grads[n_mb] = {n: p.grad.detach().clone() for n, p in model.unwrap().named_parameters()
                       if p.grad is not None}
        moe_names = {n for n, p in model.unwrap().named_parameters() if is_moe_tensor(p)}
        opt.zero_grad()
        del plugin
    assert moe_names
    for n in grads[1]:
        # two identical micro-batches -> exactly twice the single-batch gradient, for dense AND expert parameters
        torch.testing.assert_close(grads[2][n], 2 * grads[1][n], atol=1e-5, rtol=1e-4,
                                   msg=lambda m: f"{'expert' if n in moe_names else 'dense'} {n}: {m}")


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    _zero_sp()
    _pp_zero_tied()
    _moe_accum()
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_syncs_before_zero_reduction():
    spawn(_worker, 2)
