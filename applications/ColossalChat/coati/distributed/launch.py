"""Drive producers and a consumer.  In one process this alternates rollout / update (`launch_distributed(...)`); under
torchrun the ranks listed in `producer_ranks` generate while the others train, rollouts and weights moving with
`torch.distributed` object / tensor broadcasts (the reference moves them through ray object refs + a cupy NCCL group:
`coati/distributed/{launch.py, comm.py}`)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .comm import merge_rollouts
from .consumer import GRPOConsumer
from .producer import Producer


def _gather_rollouts(rollout: Optional[Dict], producer_ranks: Sequence[int], pad_token_id: int = 0) -> Dict:
    """Every rank contributes its rollout (None on consumer ranks); all ranks get the merged batch of all producers."""
    box: List[Optional[Dict]] = [None] * dist.get_world_size()
    if rollout is not None:                       # (gloo / NCCL object collectives want CPU tensors inside the pickle)
        rollout = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in rollout.items()}
    dist.all_gather_object(box, rollout)
    return merge_rollouts([box[r] for r in producer_ranks], pad_token_id)


def _broadcast_weights(sd: Optional[Dict[str, torch.Tensor]], like: Dict[str, torch.Tensor], src: int) -> Dict:
    out = {}
    for k, ref in like.items():
        t = sd[k].to(ref.device, ref.dtype).contiguous() if sd is not None else torch.empty_like(ref)
        dist.broadcast(t, src=src)
        out[k] = t
    return out


def launch_distributed(producer: Optional[Producer], consumer: Optional[GRPOConsumer], num_steps: int,
                       sync_every: int = 1, producer_ranks: Sequence[int] = (0,), consumer_src: Optional[int] = None,
                       eval_dataloaders: Optional[Dict] = None, eval_interval: int = 0, eval_reward_fn=None,
                       save_dir: Optional[str] = None, save_interval: int = 0, pad_token_id: int = 0
                       ) -> List[Dict[str, float]]:
    """Single process (no process group or world size 1): the same object owns both roles.  Multi-process: every
    rank calls this with its own role object (`producer` on producer ranks, `consumer` elsewhere); with several
    producer ranks each step trains on the merged rollouts of all of them.  `eval_interval` > 0: the (first) producer
    scores `eval_dataloaders` with `eval_reward_fn` (default: the consumer's reward) after that many updates and the
    scores land in the step's history entry; `save_interval` > 0: the consumer checkpoints into `save_dir`."""
    multi = dist.is_initialized() and dist.get_world_size() > 1
    history: List[Dict[str, float]] = []

    def after_step(step: int) -> None:
        if eval_interval and eval_dataloaders and (step + 1) % eval_interval == 0 and producer is not None \
                and (not multi or dist.get_rank() == producer_ranks[0]):
            fn = eval_reward_fn or (consumer.reward_fn if consumer is not None else None)
            scores = producer.evaluate(eval_dataloaders, fn)
            (history[-1] if history else {}).update(scores)
            producer.last_eval = scores
        if save_interval and save_dir and (step + 1) % save_interval == 0 and consumer is not None:
            consumer.save_checkpoint(save_dir)

    if not multi:
        assert producer is not None and consumer is not None
        for step in range(num_steps):
            history.append(consumer.step(producer.rollout()))
            if (step + 1) % sync_every == 0:
                producer.sync_weights(consumer.state_dict_for_producers(), consumer.version)
            after_step(step)
        return history
    rank = dist.get_rank()
    c_src = consumer_src if consumer_src is not None else min(r for r in range(dist.get_world_size())
                                                              if r not in producer_ranks)
    if consumer is not None:
        consumer.checkpoint_writer = rank == c_src
    for step in range(num_steps):
        rollout = _gather_rollouts(producer.rollout() if rank in producer_ranks else None, producer_ranks, pad_token_id)
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
        after_step(step)
    return history
