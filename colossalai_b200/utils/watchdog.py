"""Hang / failure detection for long runs.

The reference has no in-library failure detection; recovery is "torchrun restarts + resume from a
checkpoint".  This adds the missing local piece: a step heartbeat watched by a daemon thread.  When no step completes
within `timeout_s` the watchdog dumps every python thread's stack (faulthandler), optionally records the event in a
file other ranks / an external agent can poll, and aborts the process so that torchrun's `--max_restarts` can take over.
"""
from __future__ import annotations

import faulthandler
import os
import sys
import threading
import time
from typing import Callable, Optional

__all__ = ["StepWatchdog"]


class StepWatchdog:
    def __init__(self, timeout_s: float = 1800.0, on_timeout: Optional[Callable[[float], None]] = None,
                 abort: bool = True, heartbeat_file: Optional[str] = None, poll_s: float = 5.0) -> None:
        self.timeout_s, self.on_timeout, self.abort = timeout_s, on_timeout, abort
        self.heartbeat_file, self.poll_s = heartbeat_file, poll_s
        self._last = time.monotonic()
        self._step = 0
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.fired = False

    def start(self) -> "StepWatchdog":
        self._last = time.monotonic()
        self._thread = threading.Thread(target=self._run, name="cb200-watchdog", daemon=True)
        self._thread.start()
        return self

    def beat(self, step: Optional[int] = None) -> None:
        """Call once per completed training step."""
        self._last = time.monotonic()
        self._step = self._step + 1 if step is None else step
        if self.heartbeat_file:
            try:
                with open(self.heartbeat_file, "w") as f:
                    f.write(f"{self._step} {time.time():.3f}\n")
            except OSError:
                pass

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2 * self.poll_s)

    def _run(self) -> None:
        while not self._stop.wait(self.poll_s):
            idle = time.monotonic() - self._last
            if idle > self.timeout_s:
                self.fired = True
                sys.stderr.write(f"[cb200 watchdog] no step finished for {idle:.0f} s (last step {self._step}); "
                                 "dumping stacks\n")
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                if self.on_timeout is not None:
                    self.on_timeout(idle)
                if self.abort:
                    os._exit(42)
                return

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()
        return False
