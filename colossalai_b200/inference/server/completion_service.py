"""/completion service.  Parity: reference `colossalai/inference/server/completion_service.py`."""
from __future__ import annotations

from ..core.async_engine import AsyncInferenceEngine
from .utils import id_generator

__all__ = ["CompletionServing"]


class CompletionServing:
    def __init__(self, engine: AsyncInferenceEngine, served_model: str) -> None:
        self.engine = engine
        self.served_model = served_model
        try:
            import asyncio

            asyncio.get_running_loop()
        except RuntimeError:
            pass

    async def create_completion(self, request, generation_config):
        body = await request.json()
        request_id = id_generator()
        prompt = body.pop("prompt")
        final = None
        async for res in self.engine.generate(request_id, prompt, generation_config=generation_config):
            if await request.is_disconnected():
                await self.engine.abort(request_id)
                raise RuntimeError("Client disconnected")
            final = res
        return {"request_id": request_id, "model": self.served_model, "text": final}
