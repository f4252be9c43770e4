"""Per-rank event timeline -> json (-> merged plot).  Parity: reference
`colossalai/utils/rank_recorder/rank_recorder.py:27-171`."""
from __future__ import annotations

import atexit
import json
import os
import time
from typing import Dict, List

import torch


class Event:
    def __init__(self, start: float, end: float, name: str, rank: int) -> None:
        self.start, self.end, self.name, self.rank = start, end, name, rank


class Recorder:
    def __init__(self) -> None:
        self.rank_to_history: Dict[int, List[Event]] = {}
        self.base_time = time.time()
        self.temp_event = None
        self.export_format = "png"
        self.export_name = "test"
        self.dpi = 500
        self.theme = "dark_background"
        self.figure_width, self.figure_height = 20, 10
        self.legend_fontsize, self.device_fontsize, self.bar_height = 16, 20, 0.2
        self._dir = None

    def start(self, name: str = "undefined", rank: int = 0):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.temp_event = Event(time.time() - self.base_time, -1.0, name, rank)
        return self

    def end(self) -> None:
        assert self.temp_event is not None, "Recorder.end() without start()"
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        ev = self.temp_event
        ev.end = time.time() - self.base_time
        self.rank_to_history.setdefault(ev.rank, []).append(ev)
        self.temp_event = None

    def __call__(self, name: str = "undefined", rank: int = 0):
        return self.start(name, rank)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.end()

    def dump_record(self, directory: str = ".") -> None:
        os.makedirs(directory, exist_ok=True)
        for rank, hist in self.rank_to_history.items():
            with open(os.path.join(directory, f"{self.export_name}_rank{rank}.json"), "w") as f:
                json.dump({"events": [vars(e) for e in hist]}, f)

    @staticmethod
    def merge_recode(directory: str = ".", prefix: str = "test") -> Dict[int, list]:
        merged: Dict[int, list] = {}
        for fn in sorted(os.listdir(directory)):
            if fn.startswith(prefix + "_rank") and fn.endswith(".json"):
                with open(os.path.join(directory, fn)) as f:
                    data = json.load(f)
                for e in data["events"]:
                    merged.setdefault(e["rank"], []).append(e)
        return merged

    def visualize_record(self, directory: str = "."):
        import matplotlib.pyplot as plt  # optional dependency

        merged = self.merge_recode(directory, self.export_name)
        plt.figure(dpi=self.dpi, figsize=[self.figure_width, self.figure_height])
        plt.style.use(self.theme)
        for rank, evs in merged.items():
            for e in evs:
                plt.barh(y=f"device:{rank}", width=e["end"] - e["start"], left=e["start"], height=self.bar_height,
                         label=e["name"])
        plt.savefig(os.path.join(directory, f"{self.export_name}.{self.export_format}"))


recorder = Recorder()
