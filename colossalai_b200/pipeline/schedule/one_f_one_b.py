"""1F1B pipeline schedule (warm-up / steady one-forward-one-backward / cool-down).

Parity: reference `colossalai/pipeline/schedule/one_f_one_b.py:28-474`.
"""
from __future__ import annotations

from functools import partial
from typing import Any, Callable, Dict, Iterable, List, Optional, Union

import torch
from torch import Tensor
from torch.nn import Module, ModuleList
from torch.utils._pytree import tree_flatten, tree_map

from ...accelerator import get_accelerator
from ...interface import ModelWrapper, OptimizerWrapper
from ..p2p import PipelineP2PCommunication
from ..stage_manager import PipelineStageManager
from ._utils import detach, get_batch_size, get_micro_batch, merge_batch, model_forward, retain_grad, to_device
from .base import PipelineSchedule

__all__ = ["OneForwardOneBackwardSchedule"]


def default_criterion(outputs: Any, inputs: Any) -> Tensor:
    return outputs["loss"] if isinstance(outputs, dict) else outputs.loss


class OneForwardOneBackwardSchedule(PipelineSchedule):
    def __init__(self, stage_manager: PipelineStageManager, num_microbatches: Optional[int] = None,
                 microbatch_size: Optional[int] = None, enable_metadata_cache: bool = True,
                 fp8_communication: bool = False) -> None:
        super().__init__(stage_manager)
        assert num_microbatches is not None or microbatch_size is not None, (
            "Either num_microbatches or microbatch_size should be provided")
        self.comm = PipelineP2PCommunication(stage_manager, overlap_p2p=False)
        self.num_microbatch = num_microbatches
        self.microbatch_size = microbatch_size
        self.batch: Optional[Any] = None
        self.batch_size: Optional[int] = None
        self.last_batch_size: Optional[int] = None
        self.microbatch_offset: Optional[List[int]] = None
        self.enable_metadata_cache = enable_metadata_cache
        self.send_tensor_metadata = True
        self.send_grad_metadata = True
        self.tensor_metadata_recv = None
        self.grad_metadata_recv = None
        self.fp8_communication = fp8_communication

    def reset_metadata_cache(self) -> None:
        self.send_tensor_metadata = self.send_grad_metadata = True
        self.tensor_metadata_recv = self.grad_metadata_recv = None

    # ------------------------------------------------------------------ p2p wrappers with metadata caching
    def recv_forward(self) -> Any:
        if self.stage_manager.is_first_stage():
            return None
        h = self.comm._communicate(None, None, self.stage_manager.get_prev_rank(), metadata_recv=self.tensor_metadata_recv)
        out = h.wait()
        if self.enable_metadata_cache and self.tensor_metadata_recv is None:
            self.tensor_metadata_recv = h.metadata_recv
        return out

    def recv_backward(self) -> Any:
        if self.stage_manager.is_last_stage():
            return None
        h = self.comm._communicate(None, None, self.stage_manager.get_next_rank(), metadata_recv=self.grad_metadata_recv)
        out = h.wait()
        if self.enable_metadata_cache and self.grad_metadata_recv is None:
            self.grad_metadata_recv = h.metadata_recv
        return out

    def send_forward(self, output_obj: Any) -> None:
        if self.stage_manager.is_last_stage():
            return
        self.comm._communicate(output_obj, self.stage_manager.get_next_rank(), None,
                               send_metadata=self.send_tensor_metadata).wait()
        self.send_tensor_metadata = not self.enable_metadata_cache

    def send_backward(self, input_obj_grad: Any) -> None:
        if self.stage_manager.is_first_stage():
            return
        self.comm._communicate(input_obj_grad, self.stage_manager.get_prev_rank(), None,
                               send_metadata=self.send_grad_metadata).wait()
        self.send_grad_metadata = not self.enable_metadata_cache

    def send_forward_recv_backward(self, output_obj: Any, send_first: Optional[bool] = None) -> Any:
        if self.stage_manager.is_last_stage():
            return None
        h = self.comm._communicate(output_obj, self.stage_manager.get_next_rank(), self.stage_manager.get_next_rank(),
                                   send_metadata=self.send_tensor_metadata, metadata_recv=self.grad_metadata_recv,
                                   send_first=send_first)
        out = h.wait()
        self.send_tensor_metadata = not self.enable_metadata_cache
        if self.enable_metadata_cache and self.grad_metadata_recv is None:
            self.grad_metadata_recv = h.metadata_recv
        return out

    def send_backward_recv_forward(self, input_obj_grad: Any, send_first: Optional[bool] = None) -> Any:
        if self.stage_manager.is_first_stage():
            return None
        h = self.comm._communicate(input_obj_grad, self.stage_manager.get_prev_rank(),
                                   self.stage_manager.get_prev_rank(), send_metadata=self.send_grad_metadata,
                                   metadata_recv=self.tensor_metadata_recv, send_first=send_first)
        out = h.wait()
        self.send_grad_metadata = not self.enable_metadata_cache
        if self.enable_metadata_cache and self.tensor_metadata_recv is None:
            self.tensor_metadata_recv = h.metadata_recv
        return out

    # ------------------------------------------------------------------ compute steps
    def forward_step(self, model: Module, input_obj: Optional[dict], criterion: Callable,
                     accum_loss: Optional[Tensor] = None, outputs: Optional[List[Any]] = None) -> Union[Tensor, dict]:
        micro_batch = self.load_micro_batch()
        output_obj = model_forward(model, micro_batch, input_obj)
        if self.stage_manager.is_last_stage():
            loss = criterion(output_obj, micro_batch) / self.num_microbatch
            if accum_loss is not None:
                accum_loss.add_(loss.detach())
            if outputs is not None:
                outputs.append(tree_map(detach, output_obj))
            return loss
        return output_obj

    def backward_step(self, optimizer: OptimizerWrapper, input_obj: Optional[dict],
                      output_obj: Union[dict, Tensor], output_obj_grad: Optional[dict]) -> Optional[dict]:
        tree_map(retain_grad, input_obj)
        if output_obj_grad is None:
            optimizer.backward(output_obj)
        else:
            keys = output_obj.get("backward_tensor_keys", output_obj_grad.keys()) if isinstance(output_obj, dict) \
                else None
            tensors, grads = [], []
            for k, g in output_obj_grad.items():
                if isinstance(g, torch.Tensor) and isinstance(output_obj[k], torch.Tensor) and output_obj[k].requires_grad:
                    tensors.append(output_obj[k])
                    grads.append(g)
            optimizer.backward_by_grad(tensors, grads)
        input_obj_grad = None
        if input_obj is not None:
            input_obj_grad = {}
            for k, v in input_obj.items():
                if isinstance(v, torch.Tensor) and v.grad is not None:
                    input_obj_grad[k] = v.grad
        return input_obj_grad

    # ------------------------------------------------------------------ schedules
    def run_forward_only(self, model: Module, data_iter: Iterable, criterion: Callable,
                         return_loss: bool = False, return_outputs: bool = False) -> Dict:
        assert not (return_loss and not self.stage_manager.is_last_stage() and False)
        self.load_batch(data_iter)
        accum_loss = None
        if return_loss and self.stage_manager.is_last_stage():
            accum_loss = torch.scalar_tensor(0, device=get_accelerator().get_current_device())
        outputs = [] if return_outputs and self.stage_manager.is_last_stage() else None
        for _ in range(self.num_microbatch):
            input_obj = self.recv_forward()
            output_obj = self.forward_step(model, input_obj, criterion, accum_loss, outputs)
            self.send_forward(output_obj)
        if outputs is not None:
            outputs = merge_batch(outputs)
        return {"loss": accum_loss, "outputs": outputs}

    def run_forward_backward(self, model: Module, data_iter: Iterable, criterion: Callable,
                             optimizer: Optional[OptimizerWrapper] = None, return_loss: bool = False,
                             return_outputs: bool = False) -> Dict:
        self.load_batch(data_iter)
        sm = self.stage_manager
        num_warmup = min(sm.num_stages - sm.stage - 1, self.num_microbatch)
        num_remaining = self.num_microbatch - num_warmup
        input_objs, output_objs = [], []
        accum_loss = None
        if return_loss and sm.is_last_stage():
            accum_loss = torch.scalar_tensor(0, device=get_accelerator().get_current_device())
        outputs = [] if return_outputs and sm.is_last_stage() else None
        # even stages send first to break the p2p cycle (reference one_f_one_b.py:402)
        send_first = sm.stage % 2 == 0
        # ---- warm-up
        for _ in range(num_warmup):
            input_obj = self.recv_forward()
            output_obj = self.forward_step(model, input_obj, criterion, accum_loss, outputs)
            self.send_forward(output_obj)
            input_objs.append(input_obj)
            output_objs.append(output_obj)
        if num_remaining > 0:
            input_obj = self.recv_forward()
        # ---- steady 1F1B
        for i in range(num_remaining):
            output_obj = self.forward_step(model, input_obj, criterion, accum_loss, outputs)
            output_obj_grad = self.send_forward_recv_backward(output_obj, send_first=send_first)
            input_objs.append(input_obj)
            output_objs.append(output_obj)
            input_obj = input_objs.pop(0)
            output_obj = output_objs.pop(0)
            input_obj_grad = self.backward_step(optimizer, input_obj, output_obj, output_obj_grad)
            if i == num_remaining - 1:
                input_obj = None
                self.send_backward(input_obj_grad)
            else:
                input_obj = self.send_backward_recv_forward(input_obj_grad, send_first=send_first)
        # ---- cool-down
        for _ in range(num_warmup):
            input_obj = input_objs.pop(0)
            output_obj = output_objs.pop(0)
            output_obj_grad = self.recv_backward()
            input_obj_grad = self.backward_step(optimizer, input_obj, output_obj, output_obj_grad)
            self.send_backward(input_obj_grad)
        assert all(len(v) == 0 for v in (input_objs, output_objs))
        if outputs is not None:
            outputs = merge_batch(outputs)
        return {"loss": accum_loss, "outputs": outputs}

    def forward_backward_step(self, model: Module, data_iter: Iterable, criterion: Optional[Callable] = None,
                              optimizer: Optional[OptimizerWrapper] = None, return_loss: bool = False,
                              return_outputs: bool = False) -> dict:
        criterion = criterion or default_criterion
        self.forward_only = not torch.is_grad_enabled()
        if optimizer is None:
            assert self.forward_only, "Optimizer should be passed when doing backward."
        if self.forward_only:
            return self.run_forward_only(model, data_iter, criterion, return_loss, return_outputs)
        return self.run_forward_backward(model, data_iter, criterion, optimizer, return_loss, return_outputs)
