"""Graph-level intra-op auto parallelism: strategy enumeration, the ILP's plans on known-good cases (Megatron
column->row pairing, head-parallel attention, DP x TP on a 2-D mesh) and numerical equality of the transformed module
with the single-process one, forward and backward (reference: tests/test_auto_parallel/test_tensor_shard/*)."""
import copy
import math

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn

import colossalai_b200
from colossalai_b200.auto_parallel.tensor_shard import (SolverOptions, enumerate_specs, initialize_model,
                                                        reshape_dim_map, resharding_cost, resharding_steps)
from colossalai_b200.auto_parallel.tensor_shard.runtime import shard_tensor
from colossalai_b200.device import DeviceMesh
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


class Block(nn.Module):
    def __init__(self, h=64, f=256, nh=4, vocab=0):
        super().__init__()
        self.emb = nn.Embedding(vocab, h) if vocab else None
        self.ln1, self.ln2 = nn.LayerNorm(h), nn.LayerNorm(h)
        self.q, self.k, self.v, self.o = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h, bias=False)
        self.up, self.down, self.act = nn.Linear(h, f), nn.Linear(f, h), nn.GELU()
        self.gain = nn.Parameter(torch.ones(h))
        self.nh, self.hd = nh, h // nh

    def forward(self, x):
        if self.emb is not None:
            x = self.emb(x)
        B, S, H = x.shape[0], x.shape[1], x.shape[2]
        y = self.ln1(x)
        q = self.q(y).view(B, S, self.nh, self.hd).transpose(1, 2)
        k = self.k(y).view(B, S, self.nh, self.hd).transpose(1, 2)
        v = self.v(y).view(B, S, self.nh, self.hd).transpose(1, 2)
        p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.hd), dim=-1)
        c = torch.matmul(p, v).transpose(1, 2).reshape(B, S, H)
        x = x + self.o(c)
        return (x + self.down(self.act(self.up(self.ln2(x))))) * self.gain


def test_spec_enumeration_reshape_map_and_resharding():
    assert set(enumerate_specs((8, 6), (2,))) == {(None, None), (0, None), (None, 0)}
    assert (0, 1) in enumerate_specs((8, 6), (2, 2)) and (0, 0) not in enumerate_specs((8, 6), (2, 2))
    assert (None, 0) not in enumerate_specs((8, 5), (2,))                       # 5 is not divisible
    m = reshape_dim_map((2, 16, 64), (2, 16, 4, 16))                             # split heads
    assert m[2] == (2, True) and m[0] == (0, True)
    m = reshape_dim_map((2, 16, 4, 16), (2, 16, 64))                             # merge heads
    assert m[2] == (2, True) and m[3] == (2, False)
    m = reshape_dim_map((4, 16, 64), (64, 64))                                   # fold batch into tokens
    assert m[0] == (0, True) and m[1] == (0, False)
    steps = resharding_steps((0, None), (None, 0), (4,))
    assert steps == [("all_to_all", 0, 0, 1)]
    assert resharding_steps((0, None, 1), (None, None, 1), (2, 2)) == [("gather", 0, 0, None)]
    mesh = DeviceMesh(torch.arange(4), (4,))
    assert resharding_cost((0, None), (0, None), 1e6, mesh) == 0.0
    assert resharding_cost((0, None), (None, None), 1e9, mesh, False) > resharding_cost((0, None), (None, 0), 1e9, mesh, False) > 0


def test_ilp_finds_megatron_and_dp_tp_plans():
    with torch.device("meta"):
        big = Block(h=8192, f=28672, nh=64)
    mesh = DeviceMesh(torch.arange(4), (4,))
    meta = {"x": torch.empty(1, 4096, 8192, device="meta")}
    gm, sol, specs = initialize_model(big, meta, mesh, return_solution=True, apply=False)
    # column-parallel q/k/v and up, heads split through view/transpose/matmul/softmax, row-parallel o and down
    assert [sol[n] for n in ("q", "k", "v", "up")] == ["col@0"] * 4 and sol["o"] == sol["down"] == "row@0"
    assert specs["q.weight"] == (0, None) and specs["down.weight"] == (None, 0) and "down.bias" not in specs
    assert sol["softmax"] == "pointwise[R,S0,R,R]" and sol["ln1"] == "norm[R,R,R]"
    # 2-D mesh, 8 sequences: data parallel on one axis, tensor parallel on the other
    mesh2 = DeviceMesh(torch.arange(8), (2, 4))
    meta = {"x": torch.empty(8, 4096, 8192, device="meta")}
    _, sol2, specs2 = initialize_model(big, meta, mesh2, return_solution=True, apply=False)
    roles = set(sol2["up"].split("+"))
    assert any(r.startswith("col@") for r in roles) and any(r.startswith("b0@") for r in roles)
    assert sol2["x"].startswith("split[")
    tp_axis = next(a for a in specs2["up.weight"] if a is not None)
    assert specs2["down.weight"] == (None, tp_axis)
    # tiny model: latency dominates, nothing is worth sharding except free batch splits
    _, sol3, specs3 = initialize_model(Block(), {"x": torch.empty(2, 8, 64, device="meta")}, mesh,
                                       return_solution=True, apply=False)
    assert specs3 == {}
    # a memory budget below the replicated parameter size forces the weights apart
    gm3, _, _ = initialize_model(Block(), {"x": torch.empty(2, 8, 64, device="meta")}, mesh, return_solution=True,
                                 apply=False)
    total = gm3.meta["autoparallel_replicated_memory_bytes"]
    gm4, _, specs4 = initialize_model(Block(), {"x": torch.empty(2, 8, 64, device="meta")}, mesh,
                                      memory_budget=0.5 * total, return_solution=True, apply=False,
                                      solver_options=SolverOptions(shard_inputs=False))
    assert len(specs4) >= 2 and gm4.meta["autoparallel_memory_bytes"] <= 0.5 * total


def _compare(model_fn, meta, inputs, mesh, tag, **kw):
    torch.manual_seed(7)
    base = model_fn()
    ref_out = base(*inputs)
    ref_out.float().pow(2).mean().backward()
    gm, sol, specs = initialize_model(copy.deepcopy(base), meta, mesh, return_solution=True, **kw)
    for p in gm.parameters():
        p.grad = None
    out = gm(*inputs)
    torch.testing.assert_close(out, ref_out.detach(), atol=2e-5, rtol=1e-4, msg=lambda m: f"{tag}: forward {m}")
    out.float().pow(2).mean().backward()
    ref_grads = {n: p.grad for n, p in base.named_parameters()}
    n_checked = 0
    for n, p in gm.named_parameters():
        want = ref_grads[n]
        if n in specs:
            want = shard_tensor(want, specs[n], mesh)
        assert p.grad is not None, f"{tag}: {n} has no gradient"
        torch.testing.assert_close(p.grad, want, atol=2e-5, rtol=1e-4, msg=lambda m: f"{tag}: grad {n} {m}")
        n_checked += 1
    assert n_checked == len(ref_grads)
    return sol, specs


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    torch.manual_seed(11)
    x = torch.randn(2, 8, 64)
    ids = torch.randint(0, 100, (2, 8))
    meta_x = {"x": torch.empty(2, 8, 64, device="meta")}
    probe = initialize_model(Block(), meta_x, DeviceMesh(torch.arange(world_size), (world_size,)), apply=False)
    total = probe.meta["autoparallel_replicated_memory_bytes"]          # parameters + every activation, replicated
    if world_size == 2:
        mesh = DeviceMesh(torch.arange(2), (2,), init_process_group=True)
        # a) compute-bound pricing: the solver shards whatever it can (batch split here)
        fast_net = SolverOptions(peak_tflops=1e-3, hbm_gbps=1e-2)
        sol, specs = _compare(Block, meta_x, (x,), mesh, "dp", solver_options=fast_net)
        assert sol["x"].startswith("split[")
        # b) weights forced apart by the memory budget, inputs kept whole: tensor parallel
        sol, specs = _compare(Block, meta_x, (x,), mesh, "tp", memory_budget=0.6 * total,
                              solver_options=SolverOptions(shard_inputs=False))
        assert "up.weight" in specs or "down.weight" in specs, specs
        # c) embedding front end with integer inputs
        sol, specs = _compare(lambda: Block(vocab=100), {"x": torch.empty(2, 8, dtype=torch.long, device="meta")},
                              (ids,), mesh, "emb", memory_budget=0.6 * total,
                              solver_options=SolverOptions(peak_tflops=1e-3, hbm_gbps=1e-2))
    else:
        mesh = DeviceMesh(torch.arange(4), (2, 2), init_process_group=True)
        sol, specs = _compare(Block, meta_x, (x,), mesh, "2d", memory_budget=0.6 * total,
                              solver_options=SolverOptions(peak_tflops=1e-3, hbm_gbps=1e-2))
        assert len(specs) >= 2, (sol, specs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_transformed_module_matches_single_process_world2():
    spawn(_worker, 2)


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_transformed_module_matches_single_process_2d_mesh_world4():
    spawn(_worker, 4)
