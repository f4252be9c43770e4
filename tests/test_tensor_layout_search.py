"""Cost-based layout conversion (`tensor/d_tensor/layout_converter.py`): plans found by the uniform-cost search are legal,
end at the target, never cost more than the greedy heuristic under the same alpha-beta model, and executing them on a
2 x 2 gloo mesh gives exactly the target shard of the tensor - values and gradients
(reference: tests/test_tensor/test_dtensor/test_layout_converter.py)."""
import itertools

import pytest
import torch

from colossalai_b200.device import DeviceMesh as AnalyticalMesh
from colossalai_b200.tensor.d_tensor.comm_spec import CollectiveCommPattern
from colossalai_b200.tensor.d_tensor.layout import Layout
from colossalai_b200.tensor.d_tensor.layout_converter import LayoutConverter, _state_of
from colossalai_b200.tensor.d_tensor.sharding_spec import ShardingSpec
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _all_specs(ndim, n_axes):
    """Every canonical sharding of an `ndim` tensor over `n_axes` mesh axes (each axis used at most once)."""
    out = []
    for assign in itertools.product(range(-1, ndim), repeat=n_axes):      # axis -> tensor dim (or -1 = unused)
        d = {}
        for axis, dim in enumerate(assign):
            if dim >= 0:
                d.setdefault(dim, []).append(axis)
        out.append(ShardingSpec(ndim, dim_partition_dict=d))
    return out


def _replay(path, comms):
    """Check that every step is one legal transform of the state before it."""
    for before, after, cs in zip(path[:-1], path[1:], comms):
        b, a = [list(x) for x in _state_of(before.sharding_spec)], [list(x) for x in _state_of(after.sharding_spec)]
        axis = cs.logical_process_axis
        if cs.comm_pattern == CollectiveCommPattern.GATHER_FWD_SPLIT_BWD:
            assert b[cs.gather_dim][-1] == axis
            b[cs.gather_dim].pop()
        elif cs.comm_pattern == CollectiveCommPattern.SPLIT_FWD_GATHER_BWD:
            assert all(axis not in x for x in b)
            b[cs.shard_dim].append(axis)
        else:
            assert cs.comm_pattern == CollectiveCommPattern.ALL2ALL_FWD_ALL2ALL_BWD and b[cs.gather_dim][-1] == axis
            b[cs.gather_dim].pop()
            b[cs.shard_dim].append(axis)
        assert b == a, (before, after, cs)
        assert all(x == sorted(x) for x in a)                               # canonical nesting order is preserved


def test_search_plans_are_legal_and_never_worse_than_greedy():
    mesh = AnalyticalMesh(torch.arange(8), (2, 4))
    shape = torch.Size((64, 32, 16))
    specs = _all_specs(3, 2)
    conv = LayoutConverter()
    strictly_better = fewer_steps = 0
    for s, t in itertools.product(specs, specs):
        src, tgt = Layout(mesh, s, shape), Layout(mesh, t, shape)
        path, comms = conv.layout_converting(src, tgt, method="search")
        assert path[0].sharding_spec == s and path[-1].sharding_spec == t and len(comms) == len(path) - 1
        _replay(path, comms)
        gpath, gcomms = conv.layout_converting(src, tgt, method="greedy")
        _replay(gpath, gcomms)
        c_s, c_g = conv.plan_cost(src, tgt, "search"), conv.plan_cost(src, tgt, "greedy")
        assert c_s <= c_g * (1 + 1e-9) + 1e-15, (s, t, c_s, c_g)
        strictly_better += c_s < c_g * (1 - 1e-6)
        fewer_steps += len(comms) < len(gcomms)
        if s == t:
            assert comms == [] and c_s == 0.0
    assert strictly_better > 0                       # the search is not just the heuristic in disguise
    # peak memory of a plan: replicated intermediate of a gather-then-shard plan dominates
    src = Layout(mesh, ShardingSpec(3, {0: [0, 1]}), shape)
    tgt = Layout(mesh, ShardingSpec(3, {}), shape)
    path, _ = conv.layout_converting(src, tgt)
    assert conv.mem_cost(path) == pytest.approx(shape.numel() * (1 + 1 / 2))     # S01 -> S0 -> R: the last gather holds 1/2 + 1
    # forward-only planning prices local splits at zero and may pick a different plan
    fconv = LayoutConverter(forward_only=True)
    a, b = Layout(mesh, ShardingSpec(3, {}), shape), Layout(mesh, ShardingSpec(3, {1: [0, 1]}), shape)
    assert fconv.plan_cost(a, b) == 0.0 and conv.plan_cost(a, b) > 0.0


def _exec_worker(rank, world_size, port):
    import torch.distributed as dist

    import colossalai_b200
    from colossalai_b200.cluster import DeviceMesh
    from colossalai_b200.tensor.d_tensor.api import _shard_by_layout

    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    mesh = DeviceMesh(x=2, y=2)
    torch.manual_seed(0)
    full = torch.randn(8, 4, 12)
    weight = torch.arange(full.numel(), dtype=torch.float32).view_as(full) / 100
    conv = LayoutConverter()
    specs = _all_specs(3, 2)
    pairs = [(s, t) for s in specs for t in specs][:: 3]                  # a third of all pairs, both planners
    for s, t in pairs:
        for method in ("search", "greedy"):
            src, tgt = Layout(mesh, s, full.shape), Layout(mesh, t, full.shape)
            x = _shard_by_layout(full, src).clone().requires_grad_(True)
            _, comms = conv.layout_converting(src, tgt, method=method)
            y = x
            for cs in comms:
                y = cs.covert_spec_to_action(y)
            torch.testing.assert_close(y, _shard_by_layout(full, tgt), msg=lambda m: f"{s} -> {t} ({method}): {m}")
            # gradient under the tensor-parallel convention of the collectives (the backward of a gather is a local
            # slice, the backward of a split an all-gather): x.grad is this rank's SOURCE shard of the weight tensor
            if comms:
                (y * _shard_by_layout(weight, tgt)).sum().backward()
                torch.testing.assert_close(x.grad, _shard_by_layout(weight, src),
                                           msg=lambda m: f"grad {s} -> {t} ({method}): {m}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_plans_execute_to_the_target_shard_on_a_2x2_mesh():
    spawn(_exec_worker, 4)
