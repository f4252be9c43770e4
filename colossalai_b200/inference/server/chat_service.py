"""/chat service (chat templating, streaming / full responses).
Parity: reference `colossalai/inference/server/chat_service.py:14-142`."""
from __future__ import annotations

import asyncio
import codecs
import json

from ..core.async_engine import AsyncInferenceEngine
from .utils import ChatCompletionResponseStreamChoice, ChatMessage, DeltaMessage, id_generator

__all__ = ["ChatServing"]

_DEFAULT_TEMPLATE = ("{% for m in messages %}<|{{ m['role'] }}|>\n{{ m['content'] }}\n{% endfor %}"
                     "{% if add_generation_prompt %}<|assistant|>\n{% endif %}")


class ChatServing:
    def __init__(self, engine: AsyncInferenceEngine, served_model: str, tokenizer, response_role: str = "assistant",
                 chat_template=None) -> None:
        self.engine = engine
        self.served_model = served_model
        self.tokenizer = tokenizer
        self.response_role = response_role
        self._load_chat_template(chat_template)

    def _render(self, messages, add_generation_prompt: bool) -> str:
        if hasattr(self.tokenizer, "apply_chat_template") and getattr(self.tokenizer, "chat_template", None):
            return self.tokenizer.apply_chat_template(conversation=messages, tokenize=False,
                                                      add_generation_prompt=add_generation_prompt)
        out = "".join(f"<|{m['role']}|>\n{m['content']}\n" for m in messages)
        return out + (f"<|{self.response_role}|>\n" if add_generation_prompt else "")

    async def create_chat(self, request, generation_config):
        body = await request.json()
        messages = body["messages"]
        stream = str(body.pop("stream", "false")).lower()
        add_gen = body.pop("add_generation_prompt", False)
        request_id = id_generator()
        try:
            prompt = self._render(messages, add_gen)
        except Exception as e:
            raise RuntimeError(f"Error in applying chat template from request: {e}")
        result_generator = self.engine.generate(request_id, prompt, generation_config=generation_config)
        if stream == "true":
            return self.chat_completion_stream_generator(request, body, result_generator, request_id)
        return await self.chat_completion_full_generator(request, body, result_generator, request_id)

    async def chat_completion_stream_generator(self, request, request_dict, result_generator, request_id: int):
        role = self.get_chat_request_role(request, request_dict)
        first = ChatCompletionResponseStreamChoice(index=0, message=DeltaMessage(role=role))
        yield f"data: {first.model_dump_json(exclude_unset=True)}\n\n"
        async for res in result_generator:
            choice = ChatCompletionResponseStreamChoice(index=0, message=DeltaMessage(content=res))
            yield f"data: {choice.model_dump_json(exclude_unset=True)}\n\n"
        yield "data: [DONE]\n\n"

    async def chat_completion_full_generator(self, request, request_dict, result_generator, request_id):
        final = None
        async for res in result_generator:
            if await request.is_disconnected():
                await self.engine.abort(request_id)
                return {"error_msg": "Client disconnected"}
            final = res
        role = self.get_chat_request_role(request, request_dict)
        msg = ChatMessage(role=role, content=final)
        return {"request_id": request_id, "model": self.served_model,
                "choices": [{"index": 0, "message": msg.model_dump()}]}

    def get_chat_request_role(self, request, request_dict: dict) -> str:
        if not request_dict.get("add_generation_prompt", False):
            return self.response_role
        return request_dict["messages"][-1]["role"]

    def _load_chat_template(self, chat_template) -> None:
        if chat_template is None:
            return
        try:
            with open(chat_template, "r") as f:
                template = f.read()
        except OSError:
            template = codecs.decode(chat_template, "unicode_escape")
        try:
            self.tokenizer.chat_template = template
        except Exception:
            pass
