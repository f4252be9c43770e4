"""Graph capture + meta analysis.  Parity (capability, not size): reference `colossalai/fx` (`ColoTracer`,
`symbolic_trace` with `meta_args`, `ColoGraphModule`, `MetaInfoProp`, flop/memory profiler) and `colossalai/_analyzer`.
Built directly on `torch.fx`, meta tensors and `torch.utils.flop_counter`."""
from .profiler import MetaInfoProp, profile_flops_and_memory
from .passes import (activation_checkpoint_pass, balanced_split_pass, split_with_split_nodes_pass,
                     uniform_split_pass)
from .tracer import ColoGraphModule, ColoTracer, symbolic_trace

__all__ = ["ColoTracer", "ColoGraphModule", "symbolic_trace", "MetaInfoProp", "profile_flops_and_memory", "balanced_split_pass",
           "uniform_split_pass", "split_with_split_nodes_pass", "activation_checkpoint_pass"]
