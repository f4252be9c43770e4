"""Legacy sharding-spec / shape-consistency layer on the analytical device mesh (reference: tests/test_tensor/
test_sharding_spec.py, test_shape_consistency.py)."""
import pytest
import torch

from colossalai_b200.device import DeviceMesh
from colossalai_b200.tensor.shape_consistency import ShapeConsistencyManager
from colossalai_b200.tensor.sharding_spec import (DuplicatedShardingDimensionError, ShardingNotDivisibleError,
                                                  ShardingSpec)


def test_sharding_spec_and_conversion_path():
    mesh = DeviceMesh(torch.arange(8), (2, 4))
    shape = torch.Size((64, 32, 16))
    a = ShardingSpec(mesh, shape, {0: [0], 1: [1]})
    assert repr(a) == "[S0, S1, R]" and a.get_sharded_shape_per_device() == (32, 8, 16)
    b = ShardingSpec(mesh, shape, {0: [0, 1]})
    assert a.sharding_sequence_difference(b) == 2
    with pytest.raises(DuplicatedShardingDimensionError):
        ShardingSpec(mesh, shape, {0: [0], 1: [0]})
    with pytest.raises(ShardingNotDivisibleError):
        ShardingSpec(mesh, torch.Size((6, 6)), {0: [1]})
    mgr = ShapeConsistencyManager()
    path, comms, cost = mgr.shape_consistency(a, b)
    assert repr(path[0]) == repr(a) and repr(path[-1]) == repr(b) and len(comms) == len(path) - 1
    assert cost["total"] > 0 and cost["total"] == pytest.approx(cost["forward"] + cost["backward"])
    same = mgr.shape_consistency(a, a)
    assert len(same[1]) == 0 and same[2]["total"] == 0
    # the searched plan is never dearer than the heuristic one, and the peak memory of a plan is reported
    from colossalai_b200.tensor.shape_consistency import ShapeConsistencyOptions

    assert mgr.mem_cost(path) >= 2.0 * max(torch.Size(p.get_sharded_shape_per_device()).numel() for p in path)
    mgr.options = ShapeConsistencyOptions(method="greedy")
    _, _, greedy_cost = mgr.shape_consistency(a, b)
    assert cost["total"] <= greedy_cost["total"] * (1 + 1e-9)
    mgr.options = ShapeConsistencyOptions()


def _colo_spec_worker(rank, world_size, port):
    import torch.distributed as dist

    import colossalai_b200
    from colossalai_b200.legacy.tensor import (ColoTensorSpec, ComputePattern, ComputeSpec, DistSpecManager, ProcessGroup,
                                               ReplicaSpec, ShardSpec)

    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    pg = ProcessGroup(tp_degree=2, dp_degree=2)
    assert pg.tp_world_size() == 2 and pg.dp_world_size() == 2
    assert pg.tp_rank_list() == [rank - rank % 2, rank - rank % 2 + 1] and pg.dp_rank_list() == [rank % 2, rank % 2 + 2]
    assert pg == ProcessGroup(tp_degree=2, dp_degree=2) and pg.tp_local_rank() == rank % 2
    spec = ColoTensorSpec(pg, ShardSpec([0], [2]), ComputeSpec(ComputePattern.TP1D))
    assert spec.dist_attr == ShardSpec([0], [2]) and spec.dist_attr != ShardSpec([1], [2]) and ReplicaSpec() == ReplicaSpec()
    torch.manual_seed(0)
    full = torch.randn(4, 6)
    r = pg.tp_local_rank()
    rows = DistSpecManager.handle_trans_spec(full, ReplicaSpec(), ShardSpec([0], [2]), pg)
    torch.testing.assert_close(rows, full[2 * r: 2 * r + 2])
    cols = DistSpecManager.handle_trans_spec(rows, ShardSpec([0], [2]), ShardSpec([1], [2]), pg)      # all-to-all
    torch.testing.assert_close(cols, full[:, 3 * r: 3 * r + 3])
    back = DistSpecManager.handle_trans_spec(cols, ShardSpec([1], [2]), ReplicaSpec(), pg)
    torch.testing.assert_close(back, full)
    # differentiable: d/dx sum(gather(shard(x)) * w) == w
    x = full.clone().requires_grad_(True)
    w = torch.arange(24.0).view(4, 6)
    y = DistSpecManager.handle_trans_spec(x, ReplicaSpec(), ShardSpec([1], [2]), pg)
    y = DistSpecManager.handle_trans_spec(y, ShardSpec([1], [2]), ReplicaSpec(), pg)
    (y * w).sum().backward()
    torch.testing.assert_close(x.grad, w)
    with DistSpecManager.no_grad():
        z = DistSpecManager.handle_trans_spec(x, ReplicaSpec(), ShardSpec([0], [2]), pg)
    assert z.grad_fn is None or "Transform" not in type(z.grad_fn).__name__
    dist.barrier()
    dist.destroy_process_group()


def test_colo_process_group_and_dist_spec_manager():
    """reference tests/test_legacy/test_tensor/{test_dist_spec_mgr,test_process_group}.py"""
    from colossalai_b200.testing import spawn

    spawn(_colo_spec_worker, 4)
