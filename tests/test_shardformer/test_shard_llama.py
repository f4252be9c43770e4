"""Sharded-vs-single-device oracle test (reference pattern: tests/test_shardformer/test_model/test_shard_llama.py).
Runs on gloo/CPU with world_size 2 (BASELINE config 1: tiny model TP=2 plumbing)."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.cluster import DeviceMesh
from colossalai_b200.models import build_model
from colossalai_b200.shardformer import ShardConfig, ShardFormer
from colossalai_b200.shardformer.layer.utils import SeqParallelUtils
from colossalai_b200.tensor.d_tensor import is_distributed_tensor, to_global
from colossalai_b200.testing import parameterize, rerun_if_address_is_in_use, spawn

CONFIGS = [
    dict(tp=2, sp_mode=None),
    dict(tp=2, sp_mode="split_gather"),
    dict(tp=2, sp_mode="ring"),
    dict(tp=1, sp=2, sp_mode="all_to_all"),
    dict(tp=1, sp=2, sp_mode="ring_attn"),
    dict(tp=2, sp_mode=None, parallel_output=False),
]


def _run_one(model_name, cfg, atol=2e-5):
    world = dist.get_world_size()
    tp, sp = cfg.get("tp", 1), cfg.get("sp", 1)
    mesh = DeviceMesh(dp=world // (tp * sp), sp=sp, tp=tp)
    torch.manual_seed(1234)
    base = build_model(model_name)
    sharded = copy.deepcopy(base)
    sc = ShardConfig(
        tensor_parallel_process_group=mesh.group("tp"),
        sequence_parallel_process_group=mesh.group("sp") if sp > 1 else None,
        enable_tensor_parallelism=tp > 1,
        enable_sequence_parallelism=cfg["sp_mode"] is not None,
        sequence_parallelism_mode=cfg["sp_mode"],
        parallel_output=cfg.get("parallel_output", True),
    )
    sharded, _ = ShardFormer(sc).optimize(sharded)
    torch.manual_seed(7)
    ids = torch.randint(0, base.cfg.vocab_size, (2, 32))
    out_b = base(input_ids=ids, labels=ids)
    out_s = sharded(input_ids=ids, labels=ids)
    torch.testing.assert_close(out_s["loss"], out_b["loss"], atol=atol, rtol=1e-4, msg=lambda m: f"loss {model_name} {cfg}: {m}")
    out_b["loss"].backward()
    out_s["loss"].backward()
    if cfg["sp_mode"] in ("split_gather", "ring"):
        SeqParallelUtils.allreduce_partial_data_grad(mesh.group("tp"), model=sharded)
    if sp > 1:  # Ulysses / ring-attn: params replicated over sp, grads averaged over the sp group
        for p in sharded.parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad, group=mesh.group("sp"))
                p.grad /= sp
    base_grads = dict(base.named_parameters())
    checked = 0
    for name, p in sharded.named_parameters():
        if p.grad is None:
            continue
        g = p.grad
        if is_distributed_tensor(p):
            if hasattr(p, "shard_fn"):
                full = p.gather_fn(g)
            else:
                dim, group = p.dist_shard
                from colossalai_b200.parallel import comm

                full = comm.all_gather(g, dim, group)
        else:
            full = g
        ref = base_grads[name].grad
        if full.shape != ref.shape:   # padded vocab
            full = full[: ref.shape[0]]
        torch.testing.assert_close(full, ref, atol=5e-5, rtol=2e-3, msg=lambda m: f"{name} ({cfg}): {m}")
        checked += 1
    assert checked > 5
    mesh.destroy_mesh_process_groups()


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    for cfg in CONFIGS:
        _run_one("llama-tiny", cfg)
    _run_one("gpt2-tiny", dict(tp=2, sp_mode=None))
    _run_one("gpt2-tiny", dict(tp=2, sp_mode="split_gather"))
    for name in ("mixtral-tiny", "deepseek-tiny"):      # MoE blocks under pure TP / TP+SP (experts replicated)
        _run_one(name, dict(tp=2, sp_mode=None), atol=1e-4)
        _run_one(name, dict(tp=2, sp_mode="split_gather"), atol=1e-4)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_shard_llama_tp_sp():
    spawn(_worker, 2)


if __name__ == "__main__":
    test_shard_llama_tp_sp()
