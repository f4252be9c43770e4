"""Drive producers and a consumer.  In one process this alternates rollout / update (`launch_distributed(...)`); under
torchrun the ranks listed in `producer_ranks` generate while the others train, rollouts and weights moving with
`torch.distributed` object / tensor broadcasts (the reference moves them through ray object refs + a cupy NCCL group:
`coati/distributed/{launch.py, comm.py}`)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .consumer import GRPOConsumer
from .producer import Producer


def _broadcast_rollout(rollout: Optional[Dict], src: int) -> Dict:
    box = [rollout]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def _broadcast_weights(sd: Optional[Dict[str, torch.Tensor]], like: Dict[str, torch.Tensor], src: int) -> Dict:
    out = {}
    for k, ref in like.items():
        t = sd[k].to(ref.device, ref.dtype).contiguous() if sd is not None else torch.empty_like(ref)
        dist.broadcast(t, src=src)
        out[k] = t
    return out


def launch_distributed(producer: Optional[Producer], consumer: Optional[GRPOConsumer], num_steps: int,
                       sync_every: int = 1, producer_ranks: Sequence[int] = (0,), consumer_src: Optional[int] = None
                       ) -> List[Dict[str, float]]:
    """Single process (no process group or world size 1): the same object owns both roles.  Multi-process: every
    rank calls this with its own role object (`producer` on producer ranks, `consumer` elsewhere)."""
    multi = dist.is_initialized() and dist.get_world_size() > 1
    history: List[Dict[str, float]] = []
    if not multi:
        assert producer is not None and consumer is not None
        for step in range(num_steps):
            history.append(consumer.step(producer.rollout()))
            if (step + 1) % sync_every == 0:
                producer.sync_weights(consumer.state_dict_for_producers(), consumer.version)
        return history
    rank = dist.get_rank()
    p_src = producer_ranks[0]
    c_src = consumer_src if consumer_src is not None else min(r for r in range(dist.get_world_size())
                                                              if r not in producer_ranks)
    for step in range(num_steps):
        rollout = _broadcast_rollout(producer.rollout() if rank == p_src else None, p_src)
        if consumer is not None:
            history.append(consumer.step(rollout))
        if (step + 1) % sync_every == 0:
            if consumer is not None:
                like = consumer.state_dict_for_producers()
                sd = like if rank == c_src else None
            else:
                like = {k: v for k, v in producer.backend.model.state_dict().items()} \
                    if hasattr(producer.backend, "model") else producer.backend.engine.model.state_dict()
                sd = None
            new = _broadcast_weights(sd, like, c_src)
            if producer is not None:
                producer.sync_weights(new, step + 1)
    return history
