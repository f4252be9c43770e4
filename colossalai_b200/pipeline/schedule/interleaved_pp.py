"""Interleaved (virtual pipeline stages) 1F1B schedule.

Parity: reference `colossalai/pipeline/schedule/interleaved_pp.py:26-616`.  Runs on the generic node-list executor with a
node list produced by `interleaved_1f1b_schedule` (round-robin chunk placement, no dX/dW split).
"""
from __future__ import annotations

from functools import partial
from typing import Optional

from ..stage_manager import PipelineStageManager
from .v_schedule import interleaved_1f1b_schedule
from .zero_bubble_pp import NodeListScheduler

__all__ = ["InterleavedSchedule"]


class InterleavedSchedule(NodeListScheduler):
    def __init__(self, stage_manager: PipelineStageManager, num_model_chunks: int,
                 num_microbatch: Optional[int] = None, microbatch_size: Optional[int] = None,
                 enable_metadata_cache: bool = True, overlap_p2p: bool = True, fp8_communication: bool = False) -> None:
        assert num_microbatch is not None or microbatch_size is not None, (
            "Either num_microbatch or microbatch_size should be provided")
        make = partial(interleaved_1f1b_schedule, stage_manager.num_stages, n_chunk=num_model_chunks)
        super().__init__(stage_manager, lambda n_micro: make(n_micro), num_model_chunks, num_microbatch, microbatch_size,
                         v_shape=False, split_w=False, enable_metadata_cache=enable_metadata_cache,
                         overlap_p2p=overlap_p2p)
        self.fp8_communication = fp8_communication
