"""Optimizer wrappers.  Parity: reference `colossalai/interface/optimizer.py:10-187`."""
from __future__ import annotations

from typing import Dict, Optional, Union

import torch
import torch.distributed as dist
import torch.nn as nn
from torch import Tensor
from torch.optim import Optimizer


class OptimizerWrapper:
    """Standard interface around a torch optimizer: `backward`, `backward_by_grad`, clipping, grad-norm."""

    def __init__(self, optim: Optimizer) -> None:
        self.optim = optim

    @property
    def parameters(self):
        return [p for g in self.param_groups for p in g["params"]]

    @property
    def param_groups(self):
        return self.optim.param_groups

    @property
    def defaults(self):
        return self.optim.defaults

    def add_param_group(self, *args, **kwargs):
        return self.optim.add_param_group(*args, **kwargs)

    def step(self, *args, **kwargs):
        return self.optim.step(*args, **kwargs)

    def zero_grad(self, *args, **kwargs):
        self.optim.zero_grad(*args, **kwargs)

    def backward(self, loss: Tensor, inputs=None, retain_graph: bool = False, **kwargs) -> None:
        loss.backward(inputs=inputs, retain_graph=retain_graph, **kwargs)

    def backward_by_grad(self, tensor: Tensor, grad: Tensor, inputs: Tensor = None, retain_graph: bool = False):
        torch.autograd.backward(tensors=tensor, grad_tensors=grad, inputs=inputs, retain_graph=retain_graph)

    def state_dict(self):
        return self.optim.state_dict()

    def load_state_dict(self, *args, **kwargs):
        self.optim.load_state_dict(*args, **kwargs)

    def clip_grad_by_value(self, clip_value: float, *args, **kwargs) -> None:
        nn.utils.clip_grad_value_(self.parameters, clip_value, *args, **kwargs)

    def clip_grad_by_norm(self, max_norm: Union[float, int], norm_type: Union[float, int] = 2.0,
                          error_if_nonfinite: bool = False, *args, **kwargs) -> Tensor:
        return nn.utils.clip_grad_norm_(self.parameters, max_norm, norm_type, error_if_nonfinite, *args, **kwargs)

    def scale_loss(self, loss: Tensor):
        raise NotImplementedError("the method scale_loss is only available for optimizers with mixed precision")

    def unscale_grad(self):
        raise NotImplementedError("the method unscale_grad is only available for optimizers with mixed precision")

    def unwrap(self) -> Optimizer:
        return self.optim

    def get_grad_norm(self, norm_type: Union[float, int] = 2.0, **kwargs) -> Optional[float]:
        return getattr(self, "_current_grad_norm", None)


class DistributedOptim(Optimizer):
    """Interface of TP/ZeRO-aware optimizers (DistributedLamb / CAME / Adafactor / GaLore)."""

    def setup_distributed(self, tp_group: Optional[dist.ProcessGroup] = None,
                          dp_group: Optional[dist.ProcessGroup] = None,
                          shard_to_working_param: Optional[Dict] = {}, padding_map: Optional[Dict] = None,
                          is_zero: Optional[bool] = False) -> None:
        raise NotImplementedError("setup_distributed for TP/DP isn't supported by this optimizer yet!")
