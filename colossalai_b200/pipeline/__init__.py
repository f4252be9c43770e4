from .p2p import PipelineP2PCommunication
from .schedule import InterleavedSchedule, OneForwardOneBackwardSchedule, PipelineSchedule, ZeroBubbleVPipeScheduler
from .stage_manager import PipelineStageManager
from .weight_grad_store import WeightGradStore

__all__ = ["PipelineSchedule", "OneForwardOneBackwardSchedule", "InterleavedSchedule", "ZeroBubbleVPipeScheduler",
           "PipelineP2PCommunication", "PipelineStageManager", "WeightGradStore"]
