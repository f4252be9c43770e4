"""Legacy communication spec names.  Parity: reference `colossalai/tensor/comm_spec.py` (`CollectiveCommPattern`,
`CommSpec.get_comm_cost`, `covert_spec_to_action`).  Re-exports the d_tensor implementation and adds the alpha-beta cost."""
from __future__ import annotations

from typing import Dict

from .d_tensor.comm_spec import CollectiveCommPattern
from .d_tensor.comm_spec import CommSpec as _CommSpec

__all__ = ["CollectiveCommPattern", "CommSpec"]


class CommSpec(_CommSpec):
    def get_comm_cost(self, num_bytes: float = None) -> Dict[str, float]:
        """{"forward", "backward", "total"} seconds on `device_mesh` (needs an analytical `device.DeviceMesh`)."""
        mesh, axis = self.device_mesh, self.logical_process_axis
        if mesh is None or not hasattr(mesh, "all_gather_cost") or num_bytes is None:
            return {"forward": 0.0, "backward": 0.0, "total": 0.0}
        p = self.comm_pattern
        if p == CollectiveCommPattern.GATHER_FWD_SPLIT_BWD:
            f, b = mesh.all_gather_cost(num_bytes, axis), 0.0
        elif p == CollectiveCommPattern.ALL2ALL_FWD_ALL2ALL_BWD:
            f = b = mesh.all_to_all_cost(num_bytes, axis)
        elif p == CollectiveCommPattern.SPLIT_FWD_GATHER_BWD:
            f, b = 0.0, mesh.all_gather_cost(num_bytes, axis)
        elif p == CollectiveCommPattern.ALLREDUCE_FWD_IDENTITY_BWD:
            f, b = mesh.all_reduce_cost(num_bytes, axis), 0.0
        elif p == CollectiveCommPattern.IDENTITY_FWD_ALLREDUCE_BWD:
            f, b = 0.0, mesh.all_reduce_cost(num_bytes, axis)
        else:
            f = b = 0.0
        return {"forward": f, "backward": b, "total": f + b}
