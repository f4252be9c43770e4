"""device/: analytical DeviceMesh, alpha-beta profiler (gloo), pipeline-stage DP (reference: tests/test_device)."""
import numpy as np
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.device import AlphaBetaProfiler, DeviceMesh, alpa_dp, get_submesh_choices
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def test_device_mesh_shapes_and_costs():
    mesh = DeviceMesh(torch.arange(8), (2, 4))
    assert mesh.shape == (2, 4) and mesh.num_devices == 8
    assert mesh.global_rank_to_local_rank(6) == [1, 2] and mesh.global_rank_to_local_rank(6, axis=1) == 2
    assert mesh.get_ranks_in_process_group(axis=1, global_rank=5) == [4, 5, 6, 7]
    assert mesh.get_ranks_in_process_group(axis=0, global_rank=5) == [1, 5]
    assert mesh.flatten().shape == (8,)
    big, small = mesh.all_reduce_cost(1 << 30, 1), mesh.all_reduce_cost(1 << 20, 1)
    assert big > small and mesh.all_gather_cost(1 << 30, 1) < big


def test_alpa_dp_balances_stages():
    L, D = 8, 4
    choices = get_submesh_choices(1, 4)
    cost = np.full((L, L, len(choices)), np.inf)
    for i in range(L):
        for j in range(i, L):
            for m, (h, d) in enumerate(choices):
                cost[i, j, m] = (j - i + 1) / (h * d)
    total, sol = alpa_dp(L, D, 4, choices, cost)
    assert sol is not None and np.isfinite(total)
    assert sum(choices[m][0] * choices[m][1] for (_, m, _) in sol) == D
    assert sol[0][0][0] == 0 and sol[-1][0][1] == L


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    prof = AlphaBetaProfiler(list(range(world_size)), warmup=1, repeat=2, latency_iters=2)
    assert (0, 1) in prof.alpha_beta_dict and prof.alpha_beta_dict[(0, 1)][0] > 0
    mesh = prof.search_best_logical_mesh()
    assert sorted(d for row in mesh for d in row) == list(range(world_size))
    dm = DeviceMesh(torch.arange(world_size), (1, world_size), init_process_group=True)
    x = torch.ones(1)
    dist.all_reduce(x, group=dm.get_process_group(1))
    assert x.item() == world_size
    dist.barrier()
    dist.destroy_process_group()


@rerun_if_address_is_in_use()
def test_alpha_beta_profiler_gloo():
    spawn(_worker, 2)
