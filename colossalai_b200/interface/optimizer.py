"""Optimizer wrappers.  Parity: reference `colossalai/interface/optimizer.py:10-187` (`OptimizerWrapper`,
`DistributedOptim`).

`OptimizerWrapper` is the object `Booster.boost` hands back in place of the user's optimizer.  What a plugin may
override is explicit here (`backward`, `backward_by_grad`, the clipping entry points, `step` / `zero_grad`,
checkpoint state, the loss-scaling hooks of the mixed-precision subclasses); everything else an optimizer exposes -
`param_groups`, `defaults`, `state`, `add_param_group`, third-party extras - is forwarded to the wrapped optimizer by
attribute delegation instead of one pass-through method per name.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch
import torch.distributed as dist
import torch.nn as nn
from torch import Tensor
from torch.optim import Optimizer

__all__ = ["OptimizerWrapper", "DistributedOptim"]

_MIXED_PRECISION_ONLY = "the method {} is only available for optimizers with mixed precision"


class OptimizerWrapper:
    def __init__(self, optim: Optimizer) -> None:
        self.optim = optim

    # ---- everything the wrapper does not define itself comes from the wrapped optimizer
    def __getattr__(self, name: str):
        if name == "optim":                      # (not set yet: unpickling / a subclass touching attributes early)
            raise AttributeError(name)
        return getattr(self.optim, name)

    def unwrap(self) -> Optimizer:
        return self.optim

    @property
    def parameters(self) -> List[nn.Parameter]:
        return [p for group in self.optim.param_groups for p in group["params"]]

    # ---- the optimisation step
    def step(self, *args, **kwargs):
        return self.optim.step(*args, **kwargs)

    def zero_grad(self, *args, **kwargs) -> None:
        self.optim.zero_grad(*args, **kwargs)

    # ---- backward entry points (plugins route gradient synchronisation / loss scaling through these)
    def backward(self, loss: Tensor, inputs=None, retain_graph: bool = False, **kwargs) -> None:
        loss.backward(inputs=inputs, retain_graph=retain_graph, **kwargs)

    def backward_by_grad(self, tensor: Tensor, grad: Tensor, inputs: Tensor = None, retain_graph: bool = False) -> None:
        """Backward from an intermediate tensor with an incoming gradient (pipeline stages)."""
        torch.autograd.backward(tensors=tensor, grad_tensors=grad, inputs=inputs, retain_graph=retain_graph)

    # ---- gradient clipping; the sharded optimizers override these with group-aware norms
    def clip_grad_by_value(self, clip_value: float, *args, **kwargs) -> None:
        nn.utils.clip_grad_value_(self.parameters, clip_value, *args, **kwargs)

    def clip_grad_by_norm(self, max_norm: Union[float, int], norm_type: Union[float, int] = 2.0,
                          error_if_nonfinite: bool = False, *args, **kwargs) -> Tensor:
        return nn.utils.clip_grad_norm_(self.parameters, max_norm, norm_type, error_if_nonfinite, *args, **kwargs)

    def get_grad_norm(self, norm_type: Union[float, int] = 2.0, **kwargs) -> Optional[float]:
        """Global gradient norm of the last step, when the optimizer computed one (clipping enabled)."""
        return self.__dict__.get("_current_grad_norm")

    # ---- checkpoint state
    def state_dict(self):
        return self.optim.state_dict()

    def load_state_dict(self, *args, **kwargs) -> None:
        self.optim.load_state_dict(*args, **kwargs)

    # ---- loss scaling: only the mixed-precision wrappers implement these
    def scale_loss(self, loss: Tensor):
        raise NotImplementedError(_MIXED_PRECISION_ONLY.format("scale_loss"))

    def unscale_grad(self):
        raise NotImplementedError(_MIXED_PRECISION_ONLY.format("unscale_grad"))


class DistributedOptim(Optimizer):
    """Interface of the TP / ZeRO-aware optimizers (DistributedLamb / CAME / Adafactor / GaLore): the plugin calls
    `setup_distributed` once after sharding so the optimizer knows which groups to reduce its statistics over and how a
    ZeRO shard maps back to its working parameter."""

    def setup_distributed(self, tp_group: Optional[dist.ProcessGroup] = None,
                          dp_group: Optional[dist.ProcessGroup] = None,
                          shard_to_working_param: Optional[Dict] = {}, padding_map: Optional[Dict] = None,
                          is_zero: Optional[bool] = False) -> None:
        raise NotImplementedError("setup_distributed for TP/DP isn't supported by this optimizer yet!")
