"""Live-tensor leak detector.  Parity: reference `colossalai/utils/tensor_detector/tensor_detector.py:13-182`."""
from __future__ import annotations

import gc
from collections import defaultdict
from typing import Optional

import torch
import torch.nn as nn

LINE_WIDTH = 108


class TensorDetector:
    def __init__(self, show_info: bool = True, log: Optional[str] = None, include_cpu: bool = False,
                 module: Optional[nn.Module] = None) -> None:
        self.show_info, self.log, self.include_cpu = show_info, log, include_cpu
        self.module = module
        self.tensor_info = defaultdict(list)
        self.saved_tensor_info = defaultdict(list)
        self.order = []
        self.detected = []
        self.devices = []
        self.info = ""
        if module is not None:
            for name, p in module.named_parameters():
                self.tensor_info[id(p)].append(name)

    @staticmethod
    def get_tensor_mem(t: torch.Tensor) -> int:
        return t.element_size() * t.numel() + (
            t.grad.element_size() * t.grad.numel() if (t.is_leaf and t.grad is not None) else 0)

    @staticmethod
    def mem_format(b: float) -> str:
        for unit, div in (("GB", 1024**3), ("MB", 1024**2), ("KB", 1024)):
            if b >= div:
                return f"{b / div:.2f} {unit}"
        return f"{b} B"

    def collect_tensors_state(self) -> None:
        for obj in gc.get_objects():
            try:
                if not torch.is_tensor(obj):
                    continue
            except Exception:
                continue
            if not self.include_cpu and obj.device.type == "cpu":
                continue
            self.detected.append(id(obj))
            if id(obj) not in self.tensor_info:
                name = type(obj).__name__
                if isinstance(obj, nn.Parameter) and obj.grad is not None:
                    name += " (with grad)"
                self.tensor_info[id(obj)].append(name)
            info = self.tensor_info[id(obj)]
            del info[1:]
            info += [obj.device, tuple(obj.shape), obj.requires_grad, obj.dtype, self.get_tensor_mem(obj)]
            if obj.device not in self.devices:
                self.devices.append(obj.device)

    def print_tensors_state(self) -> None:
        fmt = "{:3s}{:<30s}{:>10s}{:>20s}{:>10s}{:>20s}{:>15s}"
        self.info += "\n" + "-" * LINE_WIDTH + "\n"
        self.info += fmt.format("  ", "Tensor", "device", "shape", "grad", "dtype", "Mem") + "\n" + "-" * LINE_WIDTH + "\n"
        new = [t for t in self.detected if t not in self.order]
        gone = [t for t in self.order if t not in self.detected]
        for tid in new:
            i = self.tensor_info[tid]
            self.info += fmt.format("+", str(i[0]), str(i[1]), str(i[2]), str(i[3]), str(i[4]), self.mem_format(i[5])) + "\n"
        for tid in gone:
            i = self.saved_tensor_info[tid]
            self.info += fmt.format("-", str(i[0]), str(i[1]), str(i[2]), str(i[3]), str(i[4]), self.mem_format(i[5])) + "\n"
        self.info += "-" * LINE_WIDTH + "\n"
        for dev in self.devices:
            if dev.type == "cuda":
                self.info += f"Total GPU Memory Allocated on {dev} is {self.mem_format(torch.cuda.memory_allocated(dev))}\n"
        self.info += "-" * LINE_WIDTH + "\n\n"
        if self.show_info:
            print(self.info)
        if self.log is not None:
            with open(self.log + ".log", "a") as f:
                f.write(self.info)

    def detect(self, include_cpu: Optional[bool] = None) -> None:
        if include_cpu is not None:
            self.include_cpu = include_cpu
        self.collect_tensors_state()
        self.print_tensors_state()
        self.saved_tensor_info.update(self.tensor_info)
        self.order = self.detected
        self.info = ""
        self.detected = []

    def close(self) -> None:
        self.saved_tensor_info.clear()
        self.tensor_info.clear()
        self.order, self.detected, self.devices = [], [], []
