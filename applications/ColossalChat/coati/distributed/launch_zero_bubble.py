"""Zero-bubble RL loop: the producer keeps generating into a bounded `RolloutBuffer` while the consumer trains on the
previous rollouts; fresh weights reach the producer through a latest-wins mailbox, and rollouts that fell more than
`max_staleness` policy versions behind are dropped instead of being trained on.  (The synchronous `launch_distributed`
alternates the two roles, so each one idles while the other works.)
Parity: reference `coati/distributed/launch_zero_bubble.py` + `zero_bubble/{producer,consumer,distributor}.py`."""
from __future__ import annotations

import threading
import time
from typing import Dict, List

from .comm import RolloutBuffer, WeightMailbox
from .consumer import GRPOConsumer
from .producer import Producer
from .profiling_utils import StepProfiler

__all__ = ["launch_zero_bubble"]


def launch_zero_bubble(producer: Producer, consumer: GRPOConsumer, num_steps: int, sync_every: int = 1,
                       buffer_capacity: int = 2, max_staleness: int = 2, profiler: StepProfiler = None
                       ) -> List[Dict[str, float]]:
    buf = RolloutBuffer(buffer_capacity, max_staleness)
    mail = WeightMailbox()
    prof = profiler or StepProfiler()
    errors: List[BaseException] = []

    def produce() -> None:
        try:
            while not buf.closed:
                item = mail.take()
                if item is not None:
                    producer.sync_weights(*item)
                with prof.span("rollout"):
                    r = producer.rollout()
                if not buf.push(r):
                    return
        except BaseException as e:      # surface producer failures in the consumer thread
            errors.append(e)
            buf.close()

    th = threading.Thread(target=produce, name="coati-producer", daemon=True)
    th.start()
    history: List[Dict[str, float]] = []
    try:
        for step in range(num_steps):
            with prof.span("wait_rollout"):
                rollout = buf.pop(consumer.version)
            if rollout is None:
                break
            with prof.span("train"):
                history.append(consumer.step(rollout))
            if (step + 1) % sync_every == 0:
                with prof.span("publish_weights"):
                    mail.publish(consumer.state_dict_for_producers(), consumer.version)
    finally:
        buf.close()
        th.join(timeout=60)
    if errors:
        raise errors[0]
    for h in history[-1:]:
        h.update({f"buffer_{k}": float(v) for k, v in buf.stats.items()})
    return history
