"""Schedule construction: ScheduledNode lists for interleaved 1F1B and the zero-bubble V schedule (ZB-V).

Parity: reference `colossalai/pipeline/schedule/v_schedule.py:36-449` (`ScheduledNode`, `PipelineGraph(n_stage,
n_micro, f_cost, b_cost, w_cost, c_cost, f_mem, b_mem, w_mem, max_mem).get_v_schedule()`).  The reference ports the
authors' hand-tuned search; here the schedule comes from a small event-driven list scheduler: every stage greedily
runs the highest-priority READY op (B before F before W; W fills bubbles; F is throttled by the activation-memory
limit), which reproduces the ZB-V shape (V-placed chunks, dX/dW split, deferred W) and is valid for any
(n_stage, n_micro).  Communication nodes are then inserted so that every directed channel is FIFO.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

__all__ = ["ScheduledNode", "PipelineGraph", "interleaved_1f1b_schedule", "one_f_one_b_schedule", "COMM_TYPES"]

COMM_TYPES = ("SEND_FORWARD", "RECV_FORWARD", "SEND_BACKWARD", "RECV_BACKWARD")


@dataclass(eq=True, frozen=True)
class ScheduledNode:
    type: str               # F | B | W | SEND_FORWARD | RECV_FORWARD | SEND_BACKWARD | RECV_BACKWARD
    chunk: int
    stage: int
    minibatch: int
    start_time: int = 0
    completion_time: int = 0
    rollback: bool = False


def _stage_of(chunk: int, stage: int, n_stage: int, v_shape: bool) -> int:
    """Global position (0 .. n_stage*n_chunk-1) of (chunk, stage) along the model."""
    if v_shape and chunk % 2 == 1:
        return chunk * n_stage + (n_stage - 1 - stage)
    return chunk * n_stage + stage


def _locate(pos: int, n_stage: int, v_shape: bool) -> Tuple[int, int]:
    chunk, r = divmod(pos, n_stage)
    stage = (n_stage - 1 - r) if (v_shape and chunk % 2 == 1) else r
    return chunk, stage


def _list_schedule(n_stage: int, n_micro: int, n_chunk: int, v_shape: bool, split_w: bool, f_cost: int, b_cost: int,
                   w_cost: int, c_cost: int, f_mem: float, b_mem: float, w_mem: float, max_mem: Optional[float],
                   warmup_cap: Optional[List[int]] = None) -> List[List[ScheduledNode]]:
    n_pos = n_stage * n_chunk
    done_f: Dict[Tuple[int, int], int] = {}    # (pos, mb) -> completion time
    done_b: Dict[Tuple[int, int], int] = {}
    t_stage = [0] * n_stage
    mem = [0.0] * n_stage
    pending_w: List[List[Tuple[int, int]]] = [[] for _ in range(n_stage)]
    next_f = [[0] * n_chunk for _ in range(n_stage)]
    next_b = [[0] * n_chunk for _ in range(n_stage)]
    out: List[List[ScheduledNode]] = [[] for _ in range(n_stage)]
    total_ops = n_stage * n_chunk * n_micro * (3 if split_w else 2)
    n_done = 0
    inflight = [[0] * n_chunk for _ in range(n_stage)]
    guard = 0
    while n_done < total_ops:
        guard += 1
        assert guard < 50 * total_ops + 1000, "schedule construction did not converge"
        progressed = False
        # stages take turns in order of their local clock
        for s in sorted(range(n_stage), key=lambda x: t_stage[x]):
            cands = []
            for c in range(n_chunk):
                pos = _stage_of(c, s, n_stage, v_shape)
                # backward candidate
                mb = next_b[s][c]
                if mb < n_micro and (pos, mb) in done_f:
                    ready = done_f[(pos, mb)]
                    if pos < n_pos - 1:
                        if (pos + 1, mb) not in done_b:
                            ready = None
                        else:
                            nxt_stage = _locate(pos + 1, n_stage, v_shape)[1]
                            ready = max(ready, done_b[(pos + 1, mb)] + (c_cost if nxt_stage != s else 0))
                    if ready is not None:
                        cands.append((0, max(ready, t_stage[s]), -pos, "B", c, mb))
                # forward candidate
                mb = next_f[s][c]
                if mb < n_micro:
                    ready = 0
                    ok = True
                    if pos > 0:
                        if (pos - 1, mb) not in done_f:
                            ok = False
                        else:
                            prv_stage = _locate(pos - 1, n_stage, v_shape)[1]
                            ready = done_f[(pos - 1, mb)] + (c_cost if prv_stage != s else 0)
                    if ok and max_mem is not None and mem[s] + f_mem > max_mem:
                        ok = False
                    if ok and warmup_cap is not None and inflight[s][c] >= warmup_cap[s]:
                        ok = False
                    if ok:
                        cands.append((1, max(ready, t_stage[s]), pos, "F", c, mb))
            if split_w and pending_w[s]:
                c, mb = pending_w[s][0]
                cands.append((2, t_stage[s], 0, "W", c, mb))
            if not cands:
                continue
            # earliest start first; among ops that can start now: B > F > W
            now = [x for x in cands if x[1] <= t_stage[s]]
            pick = min(now, key=lambda x: (x[0], x[2])) if now else min(cands, key=lambda x: (x[1], x[0]))
            # if the best op has to wait, fill the bubble with a W when one is pending
            if pick[1] > t_stage[s] and split_w and pending_w[s]:
                c, mb = pending_w[s][0]
                pick = (2, t_stage[s], 0, "W", c, mb)
            _, start, _, typ, c, mb = pick
            pos = _stage_of(c, s, n_stage, v_shape)
            cost = {"F": f_cost, "B": b_cost if split_w else b_cost + w_cost, "W": w_cost}[typ]
            end = start + cost
            out[s].append(ScheduledNode(typ, c, s, mb, start, end))
            t_stage[s] = end
            if typ == "F":
                done_f[(pos, mb)] = end
                next_f[s][c] += 1
                mem[s] += f_mem
                inflight[s][c] += 1
            elif typ == "B":
                done_b[(pos, mb)] = end
                next_b[s][c] += 1
                mem[s] += b_mem if split_w else (b_mem + w_mem)
                inflight[s][c] -= 1
                if split_w:
                    pending_w[s].append((c, mb))
            else:
                pending_w[s].pop(0)
                mem[s] += w_mem
            n_done += 1
            progressed = True
            break
        if not progressed:
            # everyone is blocked on memory: relax by forcing the oldest pending W / advancing time
            relaxed = False
            for s in range(n_stage):
                if pending_w[s]:
                    c, mb = pending_w[s].pop(0)
                    out[s].append(ScheduledNode("W", c, s, mb, t_stage[s], t_stage[s] + w_cost))
                    t_stage[s] += w_cost
                    mem[s] += w_mem
                    n_done += 1
                    relaxed = True
                    break
            if not relaxed:
                if max_mem is not None:
                    max_mem = None      # infeasible limit: lift it rather than dead-lock
                else:
                    raise RuntimeError("pipeline schedule dead-locked")
    return out


def _insert_comm(compute: List[List[ScheduledNode]], n_stage: int, n_chunk: int, v_shape: bool
                 ) -> List[List[ScheduledNode]]:
    """Add SEND/RECV nodes.  A send is issued right after its producer; every receive is posted right before the
    consumer.  Because both endpoints order messages of one directed channel by the PRODUCER's completion order,
    each channel (pair of ranks, direction) is FIFO."""
    n_pos = n_stage * n_chunk
    result: List[List[ScheduledNode]] = []
    for s in range(n_stage):
        nodes: List[ScheduledNode] = []
        for nd in compute[s]:
            pos = _stage_of(nd.chunk, s, n_stage, v_shape)
            if nd.type == "F" and pos > 0:
                pc, ps = _locate(pos - 1, n_stage, v_shape)
                if ps != s:
                    nodes.append(ScheduledNode("RECV_FORWARD", nd.chunk, s, nd.minibatch, nd.start_time, nd.start_time))
            if nd.type == "B" and pos < n_pos - 1:
                nc, ns = _locate(pos + 1, n_stage, v_shape)
                if ns != s:
                    nodes.append(ScheduledNode("RECV_BACKWARD", nd.chunk, s, nd.minibatch, nd.start_time, nd.start_time))
            nodes.append(nd)
            if nd.type == "F" and pos < n_pos - 1:
                nc, ns = _locate(pos + 1, n_stage, v_shape)
                if ns != s:
                    nodes.append(ScheduledNode("SEND_FORWARD", nd.chunk, s, nd.minibatch, nd.completion_time,
                                               nd.completion_time))
            if nd.type == "B" and pos > 0:
                pc, ps = _locate(pos - 1, n_stage, v_shape)
                if ps != s:
                    nodes.append(ScheduledNode("SEND_BACKWARD", nd.chunk, s, nd.minibatch, nd.completion_time,
                                               nd.completion_time))
        result.append(nodes)
    return result


class PipelineGraph:
    """ZB-V schedule builder (2 model chunks per stage, placed in a V)."""

    def __init__(self, n_stage: int, n_micro: int, f_cost: int, b_cost: int, w_cost: int, c_cost: int, f_mem: float,
                 b_mem: float, w_mem: float, max_mem: Optional[float] = None) -> None:
        self.n_node = 6 * n_stage * n_micro
        self.n_stage, self.n_micro = n_stage, n_micro
        self.f_cost, self.b_cost, self.w_cost, self.c_cost = f_cost, b_cost, w_cost, c_cost
        self.f_mem, self.b_mem, self.w_mem = f_mem, b_mem, w_mem
        self.fbw_cost = [f_cost, b_cost, w_cost]
        self.fbw_mem = [f_mem, b_mem, w_mem]
        self.max_mem = max_mem or f_mem * self.n_stage * 2

    def get_v_schedule(self, only_run_time: bool = False):
        compute = _list_schedule(self.n_stage, self.n_micro, 2, True, True, self.f_cost, self.b_cost, self.w_cost,
                                 self.c_cost, self.f_mem, self.b_mem, self.w_mem, self.max_mem)
        if only_run_time:
            return max(n[-1].completion_time for n in compute)
        return _insert_comm(compute, self.n_stage, 2, True)

    def print_details(self, schedule: List[List[ScheduledNode]]) -> str:
        lines = []
        for s, nodes in enumerate(schedule):
            lines.append(f"stage {s}: " + " ".join(f"{n.type[0]}{n.chunk}.{n.minibatch}" for n in nodes
                                                   if n.type in ("F", "B", "W")))
        return "\n".join(lines)


def interleaved_1f1b_schedule(n_stage: int, n_micro: int, n_chunk: int) -> List[List[ScheduledNode]]:
    """Interleaved (virtual-stage) 1F1B expressed as a node list (chunks placed round-robin, no dX/dW split).
    The in-flight cap reproduces Megatron's warm-up depth: (n_stage - stage - 1) * 2 + (n_chunk - 1) * n_stage + 1."""
    cap = [((n_stage - s - 1) * 2 + (n_chunk - 1) * n_stage) // n_chunk + 1 for s in range(n_stage)]
    compute = _list_schedule(n_stage, n_micro, n_chunk, False, False, 2, 2, 2, 0, 1.0, -1.0, 0.0, None,
                             warmup_cap=[max(c, 1) for c in cap])
    return _insert_comm(compute, n_stage, n_chunk, False)


def one_f_one_b_schedule(n_stage: int, n_micro: int) -> List[List[ScheduledNode]]:
    """Plain 1F1B as a node list: one chunk per stage, backward preferred over forward, at most `n_stage - stage`
    forwards in flight on a stage (its warm-up depth plus the one of the steady 1F1B pair)."""
    compute = _list_schedule(n_stage, n_micro, 1, False, False, 2, 2, 2, 0, 1.0, -1.0, 0.0, None,
                             warmup_cap=[n_stage - s for s in range(n_stage)])
    return _insert_comm(compute, n_stage, 1, False)
