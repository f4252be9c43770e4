"""Speculative decoding structs.  Parity: reference `colossalai/inference/spec/struct.py:8-56`."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

__all__ = ["DrafterOutput", "GlideInput"]


@dataclass
class DrafterOutput:
    speculated_length: int = None
    logits: torch.FloatTensor = None
    next_tokens: torch.Tensor = None
    past_key_values: Optional[Tuple[Tuple[torch.FloatTensor]]] = None

    def __post_init__(self):
        assert self.speculated_length is not None and self.speculated_length >= 0
        if self.past_key_values is not None:
            assert isinstance(self.past_key_values, tuple), "Past key values should be a tuple"


@dataclass
class GlideInput:
    """Large-model KV "glimpse" handed to a GLIDE drafter."""

    block_tables: torch.Tensor = None
    large_k_cache: torch.Tensor = None
    large_v_cache: torch.Tensor = None
    sequence_lengths: torch.Tensor = None
    n_spec_tokens: int = 5

    @property
    def glimpse_ready(self) -> bool:
        return all(a is not None for a in (self.block_tables, self.large_k_cache, self.large_v_cache,
                                           self.sequence_lengths))
