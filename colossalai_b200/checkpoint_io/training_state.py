"""Resume state beyond weights: epoch / step / sampler position / RNG streams.

The reference leaves this to applications (`applications/Colossal-LLaMA/colossal_llama/utils/ckpt_io.py:36-99`,
`StatefulDistributedSampler.set_start_index`) and does not checkpoint RNG state; here it is a small library feature."""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Iterator, Optional

import torch
import torch.distributed as dist
from torch.utils.data import DistributedSampler

__all__ = ["StatefulDistributedSampler", "save_training_state", "load_training_state"]


class StatefulDistributedSampler(DistributedSampler):
    """DistributedSampler that can resume in the middle of an epoch."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.start_index = 0

    def __iter__(self) -> Iterator:
        indices = list(super().__iter__())
        return iter(indices[self.start_index:])

    def __len__(self) -> int:
        return self.num_samples - self.start_index

    def set_start_index(self, start_index: int) -> None:
        self.start_index = start_index


def save_training_state(path: str, epoch: int, step: int, sample_start_index: int = 0,
                        extra: Optional[Dict[str, Any]] = None, save_rng: bool = True) -> None:
    """Every rank writes its RNG streams (they differ per rank); rank 0 writes the counters."""
    os.makedirs(path, exist_ok=True)
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == 0:
        with open(os.path.join(path, "running_states.json"), "w") as f:
            json.dump({"epoch": epoch, "step": step, "sample_start_index": sample_start_index, **(extra or {})}, f)
    if save_rng:
        state = {"torch": torch.get_rng_state()}
        if torch.cuda.is_available():
            state["cuda"] = torch.cuda.get_rng_state()
        try:
            import numpy as np
            import random

            state["numpy"], state["python"] = np.random.get_state(), random.getstate()
        except Exception:
            pass
        torch.save(state, os.path.join(path, f"rng_rank{rank}.pt"))


def load_training_state(path: str, sampler: Optional[StatefulDistributedSampler] = None, load_rng: bool = True
                        ) -> Dict[str, Any]:
    with open(os.path.join(path, "running_states.json")) as f:
        st = json.load(f)
    if sampler is not None:
        sampler.set_epoch(st["epoch"])
        sampler.set_start_index(st.get("sample_start_index", 0))
    rank = dist.get_rank() if dist.is_initialized() else 0
    rng_file = os.path.join(path, f"rng_rank{rank}.pt")
    if load_rng and os.path.exists(rng_file):
        state = torch.load(rng_file, weights_only=False)
        torch.set_rng_state(state["torch"])
        if "cuda" in state and torch.cuda.is_available():
            torch.cuda.set_rng_state(state["cuda"])
        if "numpy" in state:
            import numpy as np
            import random

            np.random.set_state(state["numpy"])
            random.setstate(state["python"])
    return st
