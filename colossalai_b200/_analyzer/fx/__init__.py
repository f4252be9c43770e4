from ...fx.tracer import ColoGraphModule, ColoTracer, symbolic_trace
from .node_util import MetaInfo
from .symbolic_profile import symbolic_profile

__all__ = ["symbolic_trace", "symbolic_profile", "ColoTracer", "ColoGraphModule", "MetaInfo"]
