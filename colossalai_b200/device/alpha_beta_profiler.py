"""Measure (alpha, beta) of every device pair / group and pick a logical mesh.

Parity: reference `colossalai/device/alpha_beta_profiler.py:15-393` (`profile_latency`, `profile_bandwidth`,
`profile_ab`, `search_best_logical_mesh`, `extract_alpha_beta_for_device_mesh`).  Timing uses CUDA events on GPUs (the
reference uses host wall-clock); on an NVSwitch box all pairs are homogeneous, so the search returns the trivial mesh.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["AlphaBetaProfiler"]

GB = 1 << 30
LATENCY_BYTES = 4
FRAMEWORK_LATENCY = 0


class AlphaBetaProfiler:
    def __init__(self, physical_devices: List[int], alpha_beta_dict: Optional[Dict[Tuple[int, int], Tuple[float, float]]] = None,
                 ctype: str = "a", warmup: int = 5, repeat: int = 25, latency_iters: int = 5,
                 homogeneous_tolerance: float = 0.1) -> None:
        self.physical_devices = list(physical_devices)
        self.ctype = ctype
        self.world_size = len(physical_devices)
        self.warmup, self.repeat, self.latency_iters = warmup, repeat, latency_iters
        self.homogeneous_tolerance = homogeneous_tolerance
        self.process_group_dict: Dict[Tuple[int, ...], dist.ProcessGroup] = {}
        self._init_profiling()
        self.alpha_beta_dict = alpha_beta_dict if alpha_beta_dict is not None else self.profile_ab()

    def _init_profiling(self) -> None:
        """Collective: one group per unordered device pair."""
        pairs = [(a, b) for i, a in enumerate(self.physical_devices) for b in self.physical_devices[i + 1:]]
        for p in pairs:
            self.process_group_dict[p] = dist.new_group(list(p)) if dist.is_initialized() else None

    def _device(self) -> torch.device:
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    def _profile(self, process_group: Tuple[int, ...], pg_handler, nbytes: int) -> Tuple[float, float]:
        """(seconds per op, algorithmic bandwidth B/s) of an all-reduce ('a') or broadcast over `process_group`."""
        dev = self._device()
        buf = torch.zeros(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
        me = dist.get_rank()
        if me not in process_group:
            return 0.0, 0.0

        def op():
            if self.ctype == "a":
                dist.all_reduce(buf, group=pg_handler)
            else:
                dist.broadcast(buf, src=process_group[0], group=pg_handler)

        for _ in range(self.warmup):
            op()
        if dev.type == "cuda":
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(self.repeat):
                op()
            e.record()
            torch.cuda.synchronize()
            t = s.elapsed_time(e) / 1e3 / self.repeat
        else:
            t0 = time.perf_counter()
            for _ in range(self.repeat):
                op()
            t = (time.perf_counter() - t0) / self.repeat
        n = len(process_group)
        algbw = nbytes / max(t, 1e-12)
        busbw = algbw * (2 * (n - 1) / n if self.ctype == "a" else 1.0)
        return t, busbw

    def profile_latency(self, process_group, pg_handler) -> float:
        ts = []
        for i in range(self.latency_iters):
            t, _ = self._profile(process_group, pg_handler, int(LATENCY_BYTES * (2 ** i)))
            ts.append(t)
        return sum(ts) / len(ts) if ts else 0.0

    def profile_bandwidth(self, process_group, pg_handler, maxbytes: int = GB // 4) -> float:
        _, bw = self._profile(process_group, pg_handler, maxbytes)
        return bw

    def profile_ab(self) -> Dict[Tuple[int, int], Tuple[float, float]]:
        out: Dict[Tuple[int, int], Tuple[float, float]] = {}
        dev = self._device()
        for pg, handler in self.process_group_dict.items():
            if handler is None:
                continue
            alpha = self.profile_latency(pg, handler)
            bw = self.profile_bandwidth(pg, handler, maxbytes=(GB // 4 if dev.type == "cuda" else 1 << 20))
            t = torch.tensor([alpha, bw], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)       # ranks outside the pair contribute zeros
            a, b = float(t[0]), 1.0 / max(float(t[1]), 1e-12)
            out[pg] = (a, b)
            out[(pg[1], pg[0])] = (a, b)
        return out

    # ---- mesh search
    def search_best_logical_mesh(self) -> List[List[int]]:
        """Group devices whose pairwise beta is within tolerance of the best link into rows (inner = fast axis)."""
        if self.world_size == 1 or not self.alpha_beta_dict:
            return [self.physical_devices]
        assert self.world_size & (self.world_size - 1) == 0, "the number of devices must be a power of two"
        best = min(b for (_, b) in self.alpha_beta_dict.values())
        fast = {pair for pair, (_, b) in self.alpha_beta_dict.items() if b <= best * (1 + self.homogeneous_tolerance)}
        # greedy clique growth over the "fast" graph
        remaining = list(self.physical_devices)
        rows: List[List[int]] = []
        while remaining:
            row = [remaining.pop(0)]
            for d in list(remaining):
                if all((d, r) in fast for r in row):
                    row.append(d)
                    remaining.remove(d)
            rows.append(row)
        width = min(len(r) for r in rows)
        width = 1 << int(math.log2(width))
        mesh = []
        for r in rows:
            for i in range(0, len(r) - len(r) % width, width):
                mesh.append(r[i:i + width])
        return mesh

    def extract_alpha_beta_for_device_mesh(self) -> Tuple[List[float], List[float]]:
        """(mesh_alpha, mesh_beta) for the 2-D mesh found by `search_best_logical_mesh` (axis 0 = across rows)."""
        mesh = self.search_best_logical_mesh()

        def ab(a: int, b: int) -> Tuple[float, float]:
            return self.alpha_beta_dict.get((a, b), (0.0, 0.0))

        inner = [ab(mesh[0][0], mesh[0][1])] if len(mesh[0]) > 1 else [(0.0, 0.0)]
        outer = [ab(mesh[0][0], mesh[1][0])] if len(mesh) > 1 else [(0.0, 0.0)]
        return [outer[0][0], inner[0][0]], [outer[0][1], inner[0][1]]
