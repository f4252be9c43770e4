"""Pipeline stage bookkeeping.

Parity: reference `colossalai/pipeline/stage_manager.py:11-231` (stage index from a mesh axis, wrap-around
prev/next ranks, `distribute_layers` with the remainder placed in the middle stages, interleaved / ZB-V chunk
indexing, `switch_model_chunk_id`).
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch.distributed as dist
from torch.distributed import ProcessGroup

from ..cluster import ProcessGroupMesh

__all__ = ["PipelineStageManager"]


class PipelineStageManager:
    def __init__(self, pg_mesh: ProcessGroupMesh, pipeline_axis: int, enable_interleave: bool = False,
                 use_zbv: bool = False, num_model_chunks: int = 1,
                 num_layers_per_stage: Optional[List[int]] = None) -> None:
        assert enable_interleave or num_model_chunks == 1, "num_model_chunks must be 1 when interleaving is off"
        self.pg_mesh = pg_mesh
        self.pipeline_axis = pipeline_axis
        self.num_layers_per_stage = num_layers_per_stage
        self.is_interleave = enable_interleave
        self.use_zbv = use_zbv
        self.num_model_chunks = num_model_chunks if enable_interleave else 1
        self.model_chunk_id: Optional[int] = None
        self.prev_rank: Optional[int] = None
        self.next_rank: Optional[int] = None
        self.p2p_groups: Dict[Tuple[int, ...], ProcessGroup] = {}
        self._stage = pg_mesh.coordinate(pipeline_axis)
        self._num_stages = pg_mesh.size(pipeline_axis)
        coord = pg_mesh.coordinate()
        if self._num_stages > 1:
            prev = list(coord)
            prev[pipeline_axis] = (prev[pipeline_axis] - 1) % self._num_stages
            self.prev_rank = pg_mesh.ravel(tuple(prev), pg_mesh.shape, mode="wrap")
            nxt = list(coord)
            nxt[pipeline_axis] = (nxt[pipeline_axis] + 1) % self._num_stages
            self.next_rank = pg_mesh.ravel(tuple(nxt), pg_mesh.shape, mode="wrap")
        self.pp_group = pg_mesh.get_group_along_axis(pipeline_axis) if self._num_stages > 1 else None

    # ------------------------------------------------------------------ stage queries
    @property
    def num_stages(self) -> int:
        return self._num_stages

    @property
    def stage(self) -> int:
        return self._stage

    def get_rank(self) -> int:
        return dist.get_rank()

    def is_first_stage(self, ignore_chunk: bool = False) -> bool:
        if not self.is_interleave or ignore_chunk:
            return self._stage == 0
        assert self.model_chunk_id is not None, "set model_chunk_id with switch_model_chunk_id() first"
        return self._stage == 0 and self.model_chunk_id == 0

    def is_last_stage(self, ignore_chunk: bool = False) -> bool:
        if not self.is_interleave or ignore_chunk:
            return self._stage == self._num_stages - 1
        assert self.model_chunk_id is not None, "set model_chunk_id with switch_model_chunk_id() first"
        if self.use_zbv:
            # V-shape: the last model chunk lives back on stage 0
            return self._stage == 0 and self.model_chunk_id == self.num_model_chunks - 1
        return self._stage == self._num_stages - 1 and self.model_chunk_id == self.num_model_chunks - 1

    def get_prev_rank(self) -> int:
        return self.prev_rank

    def get_next_rank(self) -> int:
        return self.next_rank

    def init_process_group_by_stages(self, stages: List[int]) -> ProcessGroup:
        return self.pg_mesh.get_group_along_axis(self.pipeline_axis, stages)

    @contextlib.contextmanager
    def switch_model_chunk_id(self, model_chunk_id: int):
        old = self.model_chunk_id
        self.model_chunk_id = model_chunk_id
        try:
            yield
        finally:
            self.model_chunk_id = old

    # ------------------------------------------------------------------ layer distribution
    def distribute_layers(self, num_layers: int, num_stages: Optional[int] = None,
                          num_model_chunks: Optional[int] = None) -> List[int]:
        """Layers per (stage x chunk) slot; when not divisible the remainder goes to the MIDDLE slots
        (first/last stages also hold embedding / head)."""
        if self.num_layers_per_stage is not None:
            assert sum(self.num_layers_per_stage) == num_layers
            return list(self.num_layers_per_stage)
        num_stages = self._num_stages if num_stages is None else num_stages
        num_model_chunks = self.num_model_chunks if num_model_chunks is None else num_model_chunks
        slots = num_stages * num_model_chunks
        quotient, remainder = divmod(num_layers, slots)
        layers = [quotient] * slots
        if remainder > 0:
            start = slots // 2 - remainder // 2
            for i in range(start, start + remainder):
                layers[i] += 1
        return layers

    def get_stage_index(self, layers_per_stage: List[int], stage: Optional[int] = None,
                        num_model_chunks: Optional[int] = None,
                        num_stages: Optional[int] = None) -> Union[Tuple[int, int], List[Tuple[int, int]]]:
        """[start, end) layer indices held by `stage`; a list of ranges when interleaved."""
        stage = self._stage if stage is None else stage
        num_model_chunks = self.num_model_chunks if num_model_chunks is None else num_model_chunks
        num_stages = self._num_stages if num_stages is None else num_stages
        cum = np.insert(np.cumsum(layers_per_stage), 0, 0)
        ranges = []
        for c in range(num_model_chunks):
            if self.use_zbv and c % 2 == 1:
                slot = c * num_stages + (num_stages - 1 - stage)   # V shape: odd chunks run backwards
            else:
                slot = c * num_stages + stage
            ranges.append((int(cum[slot]), int(cum[slot + 1])))
        if not self.is_interleave:
            return ranges[0]
        if self.model_chunk_id is not None and len(ranges) > 1:
            return ranges[self.model_chunk_id]
        return ranges
