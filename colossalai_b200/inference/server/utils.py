"""Server helpers: request ids and the chat message schemas.  Parity: reference `colossalai/inference/server/utils.py`."""
from __future__ import annotations

import itertools
import threading
from typing import Any, Optional

from pydantic import BaseModel

__all__ = ["NumericIDGenerator", "id_generator", "ChatMessage", "DeltaMessage", "ChatCompletionResponseStreamChoice"]


class NumericIDGenerator:
    """Process-wide monotonically increasing request ids.  Every construction returns the same generator, so the
    completion, chat and raw `/generate` services can never hand out the same id; `next` on the shared counter is
    taken under a lock because the services run on the event loop AND on executor threads."""

    _shared: Optional["NumericIDGenerator"] = None
    _guard = threading.Lock()

    def __new__(cls):
        with cls._guard:
            if cls._shared is None:
                inst = super().__new__(cls)
                inst._ids = itertools.count()
                cls._shared = inst
        return cls._shared

    def __call__(self) -> int:
        with self._guard:
            return next(self._ids)


id_generator = NumericIDGenerator()


class DeltaMessage(BaseModel):
    """One streamed piece of an assistant message (either field may be absent in a chunk)."""

    role: Optional[str] = None
    content: Optional[Any] = None


class ChatMessage(BaseModel):
    role: str
    content: Any


class ChatCompletionResponseStreamChoice(BaseModel):
    index: int
    message: DeltaMessage
