"""Logical KV block.  Parity: reference `colossalai/inference/kv_cache/block_cache.py`."""
from __future__ import annotations

from typing import Any

__all__ = ["CacheBlock"]


class CacheBlock:
    def __init__(self, block_id: int, block_size: int, elem_size: int, k_ptrs: Any = None, v_ptrs: Any = None) -> None:
        self.block_id = block_id
        self.block_size = block_size
        self.elem_size = elem_size
        self.k_ptrs, self.v_ptrs = k_ptrs, v_ptrs
        self.ref_count = 0
        self.allocated_size = 0
        self.token_ids = [None] * block_size

    @property
    def available_space(self) -> int:
        return self.block_size - self.allocated_size

    def add_ref(self) -> None:
        self.ref_count += 1

    def remove_ref(self) -> None:
        assert self.ref_count > 0, f"Block#{self.block_id} has no reference to remove."
        self.ref_count -= 1

    def has_ref(self) -> bool:
        return self.ref_count > 0

    def allocate(self, size: int = None) -> None:
        assert size is None or size <= self.available_space
        self.allocated_size += self.available_space if size is None else size

    def is_empty(self) -> bool:
        return self.allocated_size < 1

    def clear(self) -> None:
        self.ref_count = 0
        self.allocated_size = 0

    def __repr__(self) -> str:
        return f"CacheBlock#{self.block_id}(ref#{self.ref_count}, allocated#{self.allocated_size})"
