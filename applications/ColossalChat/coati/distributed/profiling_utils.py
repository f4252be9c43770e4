"""Wall-clock spans per role (rollout / wait / train / weight publication) with a one-line summary and an optional
append-only log file.  Parity: reference `coati/distributed/profiling_utils.py` (`CustomProfiler`)."""
from __future__ import annotations

import threading
import time
from collections import defaultdict
from contextlib import contextmanager
from typing import Dict, Optional

__all__ = ["StepProfiler"]


class StepProfiler:
    def __init__(self, name: str = "coati", log_file: Optional[str] = None) -> None:
        self.name, self.log_file = name, log_file
        self._lock = threading.Lock()
        self.total: Dict[str, float] = defaultdict(float)
        self.count: Dict[str, int] = defaultdict(int)

    @contextmanager
    def span(self, what: str):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            dt = time.perf_counter() - t0
            with self._lock:
                self.total[what] += dt
                self.count[what] += 1
            if self.log_file:
                with open(self.log_file, "a") as f:
                    f.write(f"{time.time():.3f} {self.name} {what} {dt * 1e3:.2f}ms\n")

    def summary(self) -> Dict[str, Dict[str, float]]:
        with self._lock:
            return {k: {"total_s": self.total[k], "calls": self.count[k], "mean_ms": 1e3 * self.total[k] / max(self.count[k], 1)}
                    for k in sorted(self.total)}

    def overlap_fraction(self) -> float:
        """How much of the rollout time was hidden behind training (1 = the consumer never waited for a rollout)."""
        wait, roll = self.total.get("wait_rollout", 0.0), self.total.get("rollout", 0.0)
        return 0.0 if roll <= 0 else max(0.0, 1.0 - wait / roll)
