"""Legacy pipeline tools: `PipelinableContext` (record layer construction order, partition a model into stages) and
the RPC-driven pipeline engines (`FillDrainPipelineEngine`, `OneFOneBPipelineEngine`).

Parity: reference `colossalai/legacy/pipeline/{pipelinable.py:1-260, layer_spec.py, utils.py (partition_uniform /
partition_balanced), rpc/_pipeline_base.py:1-1300, rpc/_pipeline_schedule.py:1-350, pipeline_process_group.py,
middleware/}`."""
from .pipelinable import (LayerSpec, PipelinableContext, PipelinableModel, partition_balanced, partition_uniform)
from .rpc import FillDrainPipelineEngine, OneFOneBPipelineEngine, PipelineWorker, rpc_is_initialized

__all__ = ["LayerSpec", "PipelinableContext", "PipelinableModel", "partition_uniform", "partition_balanced",
           "FillDrainPipelineEngine", "OneFOneBPipelineEngine", "PipelineWorker", "rpc_is_initialized"]
