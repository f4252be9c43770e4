"""Expert load balancer: tracks how many tokens every expert received, and periodically SWAPS experts between
expert-parallel ranks (weights, optimizer state hooks and gate rows together) so that the per-rank load evens out.

Parity: reference `colossalai/legacy/moe/load_balance.py:1-440` (`LoadBalancer.update_load`,
`_search_balance` beam search over pairwise swaps with a tolerance, `_swap_moe_param`, `balance_load`)."""
from __future__ import annotations

from copy import deepcopy
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


class LoadBalancer:
    def __init__(self, experts: nn.Module, gate: nn.Parameter, local_expert_num: int, expert_num: int, ep_group=None,
                 tolerance: float = 0.1, beam_width: int = 8, group_swap_factor: float = 0.4) -> None:
        self.experts, self.gate = experts, gate
        self.local_expert_num, self.expert_num, self.ep_group = local_expert_num, expert_num, ep_group
        self.tolerance, self.beam_width, self.group_swap_factor = tolerance, beam_width, group_swap_factor
        self.local_load: Optional[torch.Tensor] = None
        # expert placement: placement[rank][slot] = logical expert id
        n_rank = expert_num // local_expert_num
        self.placement: List[List[int]] = [list(range(r * local_expert_num, (r + 1) * local_expert_num))
                                           for r in range(n_rank)]

    # ------------------------------------------------------------------ statistics
    def update_load(self, load: torch.Tensor) -> None:
        """`load` [expert_num]: tokens routed to every (logical) expert in this step on this rank."""
        load = load.detach().float()
        self.local_load = load if self.local_load is None else self.local_load + load

    def _global_load(self) -> torch.Tensor:
        load = self.local_load.clone()
        if dist.is_initialized():
            dist.all_reduce(load)            # every data-parallel / expert-parallel rank routes different tokens
        return load

    # ------------------------------------------------------------------ search
    @staticmethod
    def _rank_loads(placement: List[List[int]], load: torch.Tensor) -> List[float]:
        return [float(sum(load[e] for e in row)) for row in placement]

    def _search_balance(self, placement: List[List[int]], load: torch.Tensor) -> Tuple[List[List[int]], List[Tuple]]:
        """Beam search over pairwise expert swaps between the most and least loaded ranks until the imbalance
        `(max - min) / mean` drops under the tolerance or no swap helps."""
        def score(p):
            rl = self._rank_loads(p, load)
            mean = sum(rl) / len(rl)
            return (max(rl) - min(rl)) / max(mean, 1e-9)

        beam = [(score(placement), placement, [])]
        best = beam[0]
        max_swaps = max(1, int(self.group_swap_factor * self.expert_num))
        for _ in range(max_swaps):
            cand = []
            for sc, p, swaps in beam:
                rl = self._rank_loads(p, load)
                hi, lo = rl.index(max(rl)), rl.index(min(rl))
                if hi == lo:
                    continue
                for i, ei in enumerate(p[hi]):
                    for j, ej in enumerate(p[lo]):
                        if load[ei] <= load[ej]:
                            continue
                        q = deepcopy(p)
                        q[hi][i], q[lo][j] = ej, ei
                        cand.append((score(q), q, swaps + [((hi, i), (lo, j))]))
            if not cand:
                break
            cand.sort(key=lambda c: c[0])
            beam = cand[: self.beam_width]
            if beam[0][0] < best[0]:
                best = beam[0]
            if best[0] <= self.tolerance:
                break
        return best[1], best[2]

    # ------------------------------------------------------------------ apply
    def _swap_moe_param(self, swaps: List[Tuple], optim=None) -> None:
        """Exchange the expert weight slices named by `swaps` between the two owning ranks (P2P over NVLink); the
        gate is indexed by LOGICAL expert id, so it does not move."""
        if not swaps:
            return
        rank = dist.get_rank(self.ep_group) if (dist.is_initialized() and self.ep_group is not None) else 0
        params = [p for p in self.experts.parameters()]
        for (ra, ia), (rb, ib) in swaps:
            if ra == rb:
                for p in params:
                    tmp = p.data[ia].clone()
                    p.data[ia] = p.data[ib]
                    p.data[ib] = tmp
                continue
            if rank not in (ra, rb):
                continue
            mine, peer = (ia, rb) if rank == ra else (ib, ra)
            peer_global = dist.get_global_rank(self.ep_group, peer)
            for p in params:
                send = p.data[mine].contiguous().clone()
                recv = torch.empty_like(send)
                ops = [dist.P2POp(dist.isend, send, peer_global), dist.P2POp(dist.irecv, recv, peer_global)]
                if rank > peer:
                    ops.reverse()
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                p.data[mine] = recv
                if optim is not None and p in getattr(optim, "state", {}):
                    for v in optim.state[p].values():        # moments of a swapped-in expert restart from zero
                        if torch.is_tensor(v) and v.shape == p.shape:
                            v[mine].zero_()

    def balance_load(self, optim=None) -> List[Tuple]:
        """Search + apply; returns the swaps performed and clears the statistics."""
        if self.local_load is None:
            return []
        load = self._global_load()
        new_placement, swaps = self._search_balance(self.placement, load)
        self._swap_moe_param(swaps, optim)
        self.placement = new_placement
        self.local_load = None
        return swaps

    def expert_location(self, expert_id: int) -> Tuple[int, int]:
        for r, row in enumerate(self.placement):
            if expert_id in row:
                return r, row.index(expert_id)
        raise KeyError(expert_id)
