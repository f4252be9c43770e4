"""Rotor: optimal activation checkpointing for a heterogeneous sequential chain under a memory budget (dynamic
program of Beaumont et al., "Optimal checkpointing for heterogeneous chains").

For a chain of stages with forward time `ftime[i]`, backward time `btime[i]`, output size `x[i+1]` and full saved-
activation size `xbar[i]`, the DP table `opt[m][i][j]` is the least time to back-propagate through stages i..j when
`m` memory slots are free and the input of stage i is resident.  Either stage i runs "save-all" (keeps xbar_i, then
the rest i+1..j), or the chain is cut at k: stages i..k-1 run forward WITHOUT saving, x_k is checkpointed, the right
part is solved first and the left part is then recomputed.

Parity: reference `colossalai/auto_parallel/checkpoint/ckpt_solver_rotor.py:1-440` (+ the C extension
`build_c_ext.py` / `dynamic_programs.c` it optionally uses — here the table is filled by a vectorised numpy DP),
`operation.py` (`Chain`, `Sequence` of `ForwardEnable / ForwardNograd / ForwardCheck / Backward` ops) and the
codegen that turns the sequence into nested `torch.utils.checkpoint` regions.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence as Seq, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint as torch_checkpoint

__all__ = ["Chain", "Sequence", "CheckpointSolverRotor", "apply_rotor_checkpointing", "profile_chain"]


@dataclass
class Chain:
    """`n` stages; `x` has n+1 entries (x[0] = chain input, x[i+1] = output of stage i)."""

    ftime: List[float]
    btime: List[float]
    x: List[float]
    xbar: List[float]
    ftmp: List[float] = field(default_factory=list)      # transient forward memory
    btmp: List[float] = field(default_factory=list)

    def __post_init__(self) -> None:
        n = len(self.ftime)
        assert len(self.btime) == n and len(self.x) == n + 1 and len(self.xbar) == n
        self.ftmp = self.ftmp or [0.0] * n
        self.btmp = self.btmp or [0.0] * n

    def __len__(self) -> int:
        return len(self.ftime)

    def discretize(self, budget: float, slots: int) -> "Chain":
        u = budget / slots
        up = lambda v: [int(math.ceil(t / u)) for t in v]
        return Chain(list(self.ftime), list(self.btime), up(self.x), up(self.xbar), up(self.ftmp), up(self.btmp))


@dataclass
class Sequence:
    """Flat op list: ("F_all", i) forward keeping everything, ("F_ck", i) forward keeping only the input (checkpoint),
    ("F_no", i) forward keeping nothing, ("B", i) backward of stage i."""

    ops: List[Tuple[str, int]] = field(default_factory=list)

    def makespan(self, chain: Chain) -> float:
        t = 0.0
        for op, i in self.ops:
            t += chain.btime[i] if op == "B" else chain.ftime[i]
        return t

    def checkpoint_segments(self) -> List[Tuple[int, int]]:
        """Top-level [start, end) stage ranges that are recomputed as one checkpointed region."""
        segs, i, ops = [], 0, self.ops
        while i < len(ops):
            op, s = ops[i]
            if op == "F_ck":
                e = s + 1
                j = i + 1
                while j < len(ops) and ops[j][0] == "F_no" and ops[j][1] == e:
                    e += 1
                    j += 1
                segs.append((s, e))
                i = j
            else:
                i += 1
        # keep only the outermost occurrence of every region (inner recomputations repeat prefixes)
        out: List[Tuple[int, int]] = []
        for s, e in segs:
            if not any(os <= s and e <= oe for os, oe in out):
                out.append((s, e))
        return sorted(out)


class CheckpointSolverRotor:
    def __init__(self, chain: Chain, memory_budget: float, memory_slots: int = 256) -> None:
        self.chain, self.budget, self.slots = chain, float(memory_budget), memory_slots
        self.d = chain.discretize(self.budget, memory_slots)
        self._opt: Optional[np.ndarray] = None
        self._what: Optional[np.ndarray] = None

    # ------------------------------------------------------------------ DP
    def _fill(self) -> None:
        c, n, M = self.d, len(self.d), self.slots
        INF = float("inf")
        opt = np.full((M + 1, n, n), INF)
        what = np.full((M + 1, n, n), -2, dtype=np.int64)          # -1: save-all first stage, k>=0: checkpoint at k
        ms = np.arange(M + 1)
        for i in range(n):
            need = max(c.x[i + 1] + c.xbar[i] + c.ftmp[i], c.x[i + 1] + c.xbar[i] + c.btmp[i])
            opt[ms >= need, i, i] = c.ftime[i] + c.btime[i]
            what[ms >= need, i, i] = -1
        for length in range(1, n):
            for i in range(n - length):
                j = i + length
                best = np.full(M + 1, INF)
                arg = np.full(M + 1, -2, dtype=np.int64)
                # (a) checkpoint x_k for k in i+1..j: forward i..k-1 without saving
                fsum = 0.0
                peak_fwd = 0
                for k in range(i + 1, j + 1):
                    fsum += c.ftime[k - 1]
                    peak_fwd = max(peak_fwd, c.x[k - 1] + c.x[k] + c.ftmp[k - 1]) if k - 1 > i else \
                        max(peak_fwd, c.x[k] + c.ftmp[k - 1])
                    mk = ms - c.x[k]
                    ok = (mk >= 0) & (ms >= peak_fwd)
                    right = np.full(M + 1, INF)
                    right[ok] = opt[mk[ok], k, j]
                    cost = fsum + right + opt[:, i, k - 1]
                    better = cost < best
                    best[better] = cost[better]
                    arg[better] = k
                # (b) save everything of stage i
                mk = ms - c.xbar[i]
                ok = (mk >= 0) & (ms >= c.x[i + 1] + c.xbar[i] + c.ftmp[i])
                right = np.full(M + 1, INF)
                right[ok] = opt[mk[ok], i + 1, j]
                cost = c.ftime[i] + right + c.btime[i]
                better = cost < best
                best[better] = cost[better]
                arg[better] = -1
                opt[:, i, j] = best
                what[:, i, j] = arg
        self._opt, self._what = opt, what

    def _rec(self, m: int, i: int, j: int, seq: List[Tuple[str, int]]) -> None:
        c = self.d
        w = int(self._what[m, i, j])
        if w == -2:
            raise RuntimeError("rotor: no feasible schedule under this budget")
        if i == j:
            seq += [("F_all", i), ("B", i)]
            return
        if w == -1:
            seq.append(("F_all", i))
            self._rec(m - c.xbar[i], i + 1, j, seq)
            seq.append(("B", i))
        else:
            k = w
            seq.append(("F_ck", i))
            seq += [("F_no", t) for t in range(i + 1, k)]
            self._rec(m - c.x[k], k, j, seq)
            self._rec(m, i, k - 1, seq)

    def solve(self) -> Sequence:
        n = len(self.chain)
        if self._opt is None:
            self._fill()
        m = self.slots - self.d.x[0]
        if m < 0 or not np.isfinite(self._opt[m, 0, n - 1]):
            raise RuntimeError(f"rotor: the chain does not fit in {self.budget:.3g} bytes even with full recomputation")
        seq: List[Tuple[str, int]] = []
        self._rec(m, 0, n - 1, seq)
        return Sequence(seq)

    def optimal_time(self) -> float:
        if self._opt is None:
            self._fill()
        return float(self._opt[self.slots - self.d.x[0], 0, len(self.chain) - 1])


# ----------------------------------------------------------------------------------------------- model glue
def profile_chain(stages: Seq[nn.Module], example: torch.Tensor, time_fn: Optional[Callable] = None) -> Chain:
    """Measure every stage of a sequential model: output bytes, saved-activation bytes (autograd saved tensors) and a
    time proxy (flops by default; pass `time_fn(stage, x) -> (fwd_s, bwd_s)` for wall-clock)."""
    from torch.utils.flop_counter import FlopCounterMode

    x = example
    ftime, btime, xs, xbar = [], [], [float(x.numel() * x.element_size())], []
    for st in stages:
        saved = [0]

        def pack(t):
            saved[0] += t.numel() * t.element_size()
            return t

        xin = x.detach().requires_grad_(x.is_floating_point())
        with torch.autograd.graph.saved_tensors_hooks(pack, lambda t: t):
            if time_fn is None:
                with FlopCounterMode(display=False) as fc:
                    y = st(xin)
                f = float(fc.get_total_flops()) or float(y.numel())
                ftime.append(f)
                btime.append(2.0 * f)
            else:
                y = st(xin)
                f, b = time_fn(st, xin)
                ftime.append(f)
                btime.append(b)
        out_b = float(y.numel() * y.element_size())
        xs.append(out_b)
        xbar.append(max(float(saved[0]), out_b))
        x = y.detach()
    return Chain(ftime, btime, xs, xbar)


class _Segment(nn.Module):
    def __init__(self, mods: Seq[nn.Module], ckpt: bool) -> None:
        super().__init__()
        self.mods, self.ckpt = nn.ModuleList(mods), ckpt

    def _run(self, x):
        for m in self.mods:
            x = m(x)
        return x

    def forward(self, x):
        if self.ckpt and self.training and torch.is_grad_enabled():
            return torch_checkpoint(self._run, x, use_reentrant=False)
        return self._run(x)


def apply_rotor_checkpointing(stages: Seq[nn.Module], example: torch.Tensor, memory_budget: float,
                              memory_slots: int = 256, time_fn: Optional[Callable] = None):
    """Solve for `stages` and return (`nn.Sequential` with the chosen segments wrapped in activation checkpoints,
    the `Sequence`, the profiled `Chain`)."""
    chain = profile_chain(stages, example, time_fn)
    seq = CheckpointSolverRotor(chain, memory_budget, memory_slots).solve()
    segs = seq.checkpoint_segments()
    out: List[nn.Module] = []
    i = 0
    stages = list(stages)
    for s, e in segs:
        if s > i:
            out.append(_Segment(stages[i:s], ckpt=False))
        out.append(_Segment(stages[s:e], ckpt=True))
        i = e
    if i < len(stages):
        out.append(_Segment(stages[i:], ckpt=False))
    return nn.Sequential(*out), seq, chain
