"""Logical KV block: bookkeeping of one physical block of the paged cache (who references it, how many token slots
are filled).  Parity: reference `colossalai/inference/kv_cache/block_cache.py`."""
from __future__ import annotations

from typing import Any, List, Optional

__all__ = ["CacheBlock"]


class CacheBlock:
    __slots__ = ("block_id", "block_size", "elem_size", "k_ptrs", "v_ptrs", "ref_count", "allocated_size", "token_ids")

    def __init__(self, block_id: int, block_size: int, elem_size: int, k_ptrs: Any = None, v_ptrs: Any = None) -> None:
        self.block_id, self.block_size, self.elem_size = block_id, block_size, elem_size
        self.k_ptrs, self.v_ptrs = k_ptrs, v_ptrs
        self.token_ids: List[Optional[int]] = [None] * block_size
        self.clear()

    def clear(self) -> None:
        """Back to the free state: nobody references the block, no slot is filled."""
        self.ref_count = 0
        self.allocated_size = 0

    # ---- references (a block shared by several sequences - beam search, prefix sharing - is freed with the last one)
    def add_ref(self) -> None:
        self.ref_count += 1

    def remove_ref(self) -> None:
        if self.ref_count <= 0:
            raise AssertionError(f"Block#{self.block_id} has no reference to remove.")
        self.ref_count -= 1

    def has_ref(self) -> bool:
        return self.ref_count > 0

    # ---- token slots
    @property
    def available_space(self) -> int:
        return self.block_size - self.allocated_size

    def allocate(self, size: Optional[int] = None) -> None:
        """Fill `size` more slots (all remaining ones when omitted)."""
        room = self.available_space
        size = room if size is None else size
        assert size <= room, f"Block#{self.block_id}: {size} slots requested, {room} left"
        self.allocated_size += size

    def is_empty(self) -> bool:
        return self.allocated_size == 0

    def __repr__(self) -> str:
        return f"CacheBlock#{self.block_id}(ref#{self.ref_count}, allocated#{self.allocated_size})"
