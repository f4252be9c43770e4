"""multi_tensor_applier: run one fused native op over lists of tensors.
Parity: reference `colossalai/utils/multi_tensor_apply/multi_tensor_apply.py` (chunk size 2048*32)."""
from __future__ import annotations


class MultiTensorApply:
    def __init__(self, chunk_size: int = 2048 * 32) -> None:
        self.chunk_size = chunk_size

    def __call__(self, op, noop_flag_buffer, tensor_lists, *args):
        for i, l in enumerate(tensor_lists):
            assert isinstance(l, (list, tuple)), f"tensor_lists[{i}] must be a list"
        return op(self.chunk_size, noop_flag_buffer, tensor_lists, *args)


multi_tensor_applier = MultiTensorApply(2048 * 32)
