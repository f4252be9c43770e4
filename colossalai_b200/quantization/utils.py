"""FSDP parameter-communication hook plumbing.  Parity: reference `colossalai/quantization/utils.py:9-112`
(`register_params_comm_hook`, `patch_fsdp_params_comm_hook`)."""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["register_params_comm_hook", "patch_fsdp_params_comm_hook"]


def _all_gather_flat_param(self, padded_unsharded_flat_param: torch.Tensor) -> torch.Tensor:
    """Replacement of `FlatParamHandle._all_gather_flat_param` that routes through a registered hook."""
    sharded = self.flat_param.data
    expected = sharded.numel() * self.world_size
    assert padded_unsharded_flat_param.numel() == expected
    pg = self._fake_process_group if getattr(self, "_use_fake_all_gather", False) else self.process_group
    hook = getattr(self, "_comm_hook", None)
    if hook is not None:
        hook(getattr(self, "_comm_hook_state", None), padded_unsharded_flat_param, sharded, pg)
    elif sharded.is_cpu:
        chunks = list(padded_unsharded_flat_param.chunk(dist.get_world_size(pg)))
        dist.all_gather(chunks, sharded, group=pg)
    else:
        dist.all_gather_into_tensor(padded_unsharded_flat_param, sharded, pg)
    return padded_unsharded_flat_param


def register_params_comm_hook(self, state: object, hook: callable) -> None:
    """Bound onto FSDP: registers `hook(state, padded_unsharded_flat_param, sharded_flat_param, group)`."""
    if not self.check_is_root():
        raise AssertionError("register_comm_hook can only be called on a root instance.")
    import torch.distributed.fsdp._traversal_utils as traversal_utils

    for fsdp_state in traversal_utils._get_fsdp_states(self):
        h = getattr(fsdp_state, "_handle", None)
        if h is not None:
            assert getattr(h, "_comm_hook", None) is None, "A communication hook is already registered"
            if not callable(hook):
                raise ValueError(f"The communication hook must be callable but got {hook}")
            h._comm_hook, h._comm_hook_state = hook, state


def patch_fsdp_params_comm_hook() -> None:
    from torch.distributed.fsdp import FullyShardedDataParallel as FSDP
    from torch.distributed.fsdp._flat_param import FlatParamHandle

    FlatParamHandle._comm_hook = None
    FlatParamHandle._comm_hook_state = None
    FlatParamHandle._all_gather_flat_param = _all_gather_flat_param
    FSDP.register_params_comm_hook = register_params_comm_hook
