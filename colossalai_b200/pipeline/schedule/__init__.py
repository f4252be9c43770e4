from .base import PipelineSchedule
from .interleaved_pp import InterleavedSchedule
from .one_f_one_b import OneForwardOneBackwardSchedule
from .v_schedule import PipelineGraph, ScheduledNode
from .zero_bubble_pp import NodeListScheduler, ZeroBubbleVPipeScheduler

__all__ = ["PipelineSchedule", "OneForwardOneBackwardSchedule", "InterleavedSchedule", "ZeroBubbleVPipeScheduler",
           "NodeListScheduler", "PipelineGraph", "ScheduledNode", "GenerateSchedule"]


def __getattr__(name):
    if name == "GenerateSchedule":
        from .generate import GenerateSchedule

        return GenerateSchedule
    raise AttributeError(name)
