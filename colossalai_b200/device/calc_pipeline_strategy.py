"""Alpa-style inter-op dynamic program: split a layer sequence into pipeline stages over sub-meshes.
Parity: reference `colossalai/device/calc_pipeline_strategy.py:6-131` (`get_submesh_choices`, `alpa_dp`)."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

__all__ = ["get_submesh_choices", "alpa_dp", "alpa_dp_impl"]


def get_submesh_choices(num_hosts: int, num_devices_per_host: int, mode: str = "new") -> List[Tuple[int, int]]:
    choices = []
    i = 1
    while i <= num_devices_per_host:          # power-of-two slices inside a host
        choices.append((1, i))
        i *= 2
    assert choices[-1][1] == num_devices_per_host, "num_devices_per_host must be a power of two"
    if mode == "alpa":
        for h in range(2, num_hosts + 1):
            choices.append((h, num_devices_per_host))
    else:
        h = 2
        while h <= num_hosts:
            choices.append((h, num_devices_per_host))
            h *= 2
    return choices


def alpa_dp_impl(num_layers: int, num_devices: int, num_microbatches: int, submesh_choices: Sequence[Tuple[int, int]],
                 compute_cost: np.ndarray, max_stage_cost: float, best_configs: np.ndarray):
    """f[s, l, d] = min total cost of running layers l.. on d devices with s stages, every stage <= max_stage_cost."""
    inf = float("inf")
    f = np.full((num_layers + 1, num_layers + 1, num_devices + 1), inf, dtype=np.float64)
    f_stage_max = np.zeros_like(f)
    f_arg = np.full((num_layers + 1, num_layers + 1, num_devices + 1, 3), -1, dtype=np.int64)
    f[0, num_layers, 0] = 0.0
    for s in range(1, num_layers + 1):
        for i in range(num_layers - 1, -1, -1):
            for j in range(1, num_devices + 1):
                for k in range(num_layers, i, -1):
                    for m, sub in enumerate(submesh_choices):
                        n_sub = sub[0] * sub[1]
                        if n_sub > j:
                            continue
                        stage_cost = compute_cost[i, k - 1, m]
                        if stage_cost > max_stage_cost:
                            continue
                        new = f[s - 1, k, j - n_sub] + stage_cost
                        if new < f[s, i, j]:
                            f[s, i, j] = new
                            f_stage_max[s, i, j] = max(f_stage_max[s - 1, k, j - n_sub], stage_cost)
                            f_arg[s, i, j] = (k, m, best_configs[i, k - 1, m])
    best_s, best_total = -1, inf
    for s in range(1, num_layers + 1):
        if f[s, 0, num_devices] < inf:
            total = f[s, 0, num_devices] + (num_microbatches - 1) * f_stage_max[s, 0, num_devices]
            if total < best_total:
                best_total, best_s = total, s
    if best_s < 0:
        return inf, None
    res, i, j, s = [], 0, num_devices, best_s
    while s > 0 and i < num_layers and j > 0:
        k, m, cfg = f_arg[s, i, j]
        res.append(((i, int(k)), int(m), int(cfg)))
        j -= submesh_choices[int(m)][0] * submesh_choices[int(m)][1]
        i, s = int(k), s - 1
    return best_total, res


def alpa_dp(num_layers: int, num_devices: int, num_microbatches: int, submesh_choices: Sequence[Tuple[int, int]],
            compute_cost: np.ndarray, best_configs: np.ndarray = None, gap: float = 1e-6):
    """Try every distinct stage cost as the max-stage bound (ascending) and keep the best total latency."""
    if best_configs is None:
        best_configs = np.zeros(compute_cost.shape, dtype=np.int64)
    costs = np.sort(np.unique(compute_cost[np.isfinite(compute_cost)]))
    best_cost, best_sol, last = float("inf"), None, 0.0
    for c in costs:
        if c * num_microbatches >= best_cost:
            break
        if c - last < gap:
            continue
        cost, sol = alpa_dp_impl(num_layers, num_devices, num_microbatches, submesh_choices, compute_cost, c,
                                 best_configs)
        if sol is not None and cost < best_cost:
            best_cost, best_sol = cost, sol
        last = c
    return best_cost, best_sol
