"""Server helpers.  Parity: reference `colossalai/inference/server/utils.py`."""
from __future__ import annotations

import itertools
from typing import Any, Optional

from pydantic import BaseModel

__all__ = ["NumericIDGenerator", "id_generator", "ChatMessage", "DeltaMessage", "ChatCompletionResponseStreamChoice"]


class NumericIDGenerator:
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
            cls._instance._counter = itertools.count()
        return cls._instance

    def __call__(self) -> int:
        return next(self._counter)


id_generator = NumericIDGenerator()


class ChatMessage(BaseModel):
    role: str
    content: Any


class DeltaMessage(BaseModel):
    role: Optional[str] = None
    content: Optional[Any] = None


class ChatCompletionResponseStreamChoice(BaseModel):
    index: int
    message: DeltaMessage
