"""1F1B pipeline schedule (warm-up / steady one-forward-one-backward / cool-down).

Parity: reference `colossalai/pipeline/schedule/one_f_one_b.py:28-474` (same constructor and `forward_backward_step`
contract).  There is no hand-written warm-up / steady / cool-down loop here: 1F1B is the single-chunk case of the
node-list executor (`zero_bubble_pp.NodeListScheduler`) - `one_f_one_b_schedule` emits, per stage, the F / B order with
at most `n_stage - stage` forwards in flight, and the executor runs it over the directed per-pair channels (asynchronous
sends, receives posted in the sender's order), exactly like the interleaved and zero-bubble schedules.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

from ..stage_manager import PipelineStageManager
from ._utils import default_criterion  # noqa: F401  (re-exported: older call sites import it from here)
from .v_schedule import one_f_one_b_schedule
from .zero_bubble_pp import NodeListScheduler

__all__ = ["OneForwardOneBackwardSchedule"]


class OneForwardOneBackwardSchedule(NodeListScheduler):
    def __init__(self, stage_manager: PipelineStageManager, num_microbatches: Optional[int] = None,
                 microbatch_size: Optional[int] = None, enable_metadata_cache: bool = True,
                 fp8_communication: bool = False) -> None:
        assert num_microbatches is not None or microbatch_size is not None, (
            "Either num_microbatches or microbatch_size should be provided")
        super().__init__(stage_manager, partial(one_f_one_b_schedule, stage_manager.num_stages), 1, num_microbatches,
                         microbatch_size, v_shape=False, split_w=False, enable_metadata_cache=enable_metadata_cache)
        self.fp8_communication = fp8_communication
