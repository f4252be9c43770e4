"""Process-wide MoE settings + auxiliary-loss accumulator.  Parity: reference `legacy/moe/manager.py:1-160`."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from ...context import SingletonMeta


class MoEManager(metaclass=SingletonMeta):
    def __init__(self) -> None:
        self.parallel: Optional[str] = None       # None | "EP"
        self.ep_size = 1
        self.ep_group = None
        self.use_kernel_optim = True
        self.has_setup = False
        self._aux_loss: List[torch.Tensor] = []
        self._z_loss: List[torch.Tensor] = []

    def setup(self, parallel: Optional[str] = None, ep_size: Optional[int] = None, ep_group=None,
              use_kernel_optim: bool = True, **unused) -> None:
        self.parallel = parallel
        self.use_kernel_optim = use_kernel_optim
        if parallel == "EP":
            if ep_group is None and dist.is_initialized():
                world = dist.get_world_size()
                ep_size = ep_size or world
                assert world % ep_size == 0
                rank = dist.get_rank()
                for start in range(0, world, ep_size):
                    g = dist.new_group(list(range(start, start + ep_size)))
                    if start <= rank < start + ep_size:
                        ep_group = g
            self.ep_group = ep_group
            self.ep_size = dist.get_world_size(ep_group) if (ep_group is not None and dist.is_initialized()) else 1
        else:
            self.ep_group, self.ep_size = None, 1
        self.has_setup = True

    # auxiliary losses are appended by routers during forward and summed by the training loop
    def reset_loss(self) -> None:
        self._aux_loss, self._z_loss = [], []

    def add_loss(self, aux_loss: float = 0.0, z_loss: float = 0.0) -> None:
        self._aux_loss.append(aux_loss)
        self._z_loss.append(z_loss)

    def get_loss(self):
        return self._aux_loss, self._z_loss


MOE_MANAGER = MoEManager()
