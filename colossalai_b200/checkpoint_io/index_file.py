"""HF-style checkpoint index (`*.index.json`: metadata + weight_map).
Parity: reference `colossalai/checkpoint_io/index_file.py:12-182`."""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from pathlib import Path
from typing import Any, Dict, List, Union

__all__ = ["CheckpointIndexFile"]


class CheckpointIndexFile:
    def __init__(self, root_path: Union[str, Path, None] = None) -> None:
        self.root_path = root_path
        self.metadata: Dict[str, Any] = {}
        self.weight_map: Dict[str, str] = OrderedDict()

    @staticmethod
    def from_file(index_path: Union[str, Path]) -> "CheckpointIndexFile":
        index = CheckpointIndexFile()
        index.load(index_path)
        return index

    def load(self, json_path: Union[str, Path]) -> None:
        with open(json_path, "r", encoding="utf-8") as f:
            data = json.load(f)
        self.metadata = data.get("metadata", {})
        self.weight_map = OrderedDict(data.get("weight_map", {}))
        for k, v in data.items():
            if k not in ("metadata", "weight_map"):
                self.metadata[k] = v
        self.root_path = Path(json_path).absolute().parent

    def export(self, json_path: Union[str, Path]) -> None:
        with open(json_path, "w", encoding="utf-8") as f:
            json.dump({"metadata": self.metadata, "weight_map": self.weight_map}, f, indent=2, sort_keys=True)

    def write_index_file(self, save_index_file: str) -> None:
        self.export(os.path.join(self.root_path, save_index_file))

    def append_weight_map(self, param_name: str, shard_file: str) -> None:
        self.weight_map[param_name] = shard_file

    def append_meta_data(self, name: str, val: Any) -> None:
        self.metadata[name] = val

    def contains_dtensor(self) -> bool:
        return any(v.endswith(".*.bin") or v.endswith(".*.safetensors") for v in self.weight_map.values())

    def get_checkpoint_filenames(self) -> List[str]:
        seen, out = set(), []
        for v in self.weight_map.values():
            if v not in seen:
                seen.add(v)
                out.append(str(Path(self.root_path) / v))
        return out

    def assert_no_dtensor_checkpoint(self) -> None:
        assert not self.contains_dtensor(), "checkpoint contains distributed-tensor shards"

    def get_checkpoint_file(self, param_name: str) -> str:
        return self.weight_map[param_name]

    def get_all_param_names(self) -> List[str]:
        return list(self.weight_map.keys())

    def get_param_group_filename(self):
        fn = self.metadata.get("param_groups")
        return None if fn is None else str(Path(self.root_path) / fn)
