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
