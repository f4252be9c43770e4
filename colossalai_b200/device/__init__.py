from .alpha_beta_profiler import AlphaBetaProfiler
from .calc_pipeline_strategy import alpa_dp, get_submesh_choices
from .device_mesh import DeviceMesh

__all__ = ["DeviceMesh", "AlphaBetaProfiler", "alpa_dp", "get_submesh_choices"]
