"""FastAPI HTTP server: /ping /engine_check /generate /completion /chat.

    python -m colossalai_b200.inference.server.api_server --model llama-tiny --port 8000

Parity: reference `colossalai/inference/server/api_server.py:1-237`."""

import argparse
import json

from ..config import GenerationConfig, InferenceConfig
from ..core.async_engine import AsyncInferenceEngine
from .chat_service import ChatServing
from .completion_service import CompletionServing
from .utils import id_generator

TIMEOUT_KEEP_ALIVE = 5
prompt_template_choices = ["llama", "vicuna"]


def get_generation_config(request: dict) -> GenerationConfig:
    cfg = GenerationConfig()
    for k, v in request.items():
        if hasattr(cfg, k):
            setattr(cfg, k, v)
    return cfg


def build_app(async_engine: AsyncInferenceEngine, served_model: str = "model", tokenizer=None,
              response_role: str = "assistant", chat_template=None):
    from fastapi import FastAPI, Request
    from fastapi.responses import JSONResponse, Response, StreamingResponse

    app = FastAPI()
    tok = tokenizer or async_engine.engine.engine.tokenizer
    completion_serving = CompletionServing(async_engine, served_model)
    chat_serving = ChatServing(async_engine, served_model, tok, response_role, chat_template)

    @app.get("/ping")
    def health_check():
        return JSONResponse({"status": "Healthy"})

    @app.get("/metrics")
    def metrics():
        """Prometheus text exposition of the scheduler state (requests waiting / running, KV blocks in use)."""
        rh = async_engine.engine.engine.request_handler
        cm = rh.cache_manager
        waiting = sum(len(l) for l in rh.waiting_list)
        running = rh.total_requests_in_batch_bucket()
        lines = ["# TYPE cb200_requests_waiting gauge", f"cb200_requests_waiting {waiting}",
                 "# TYPE cb200_requests_running gauge", f"cb200_requests_running {running}",
                 "# TYPE cb200_kv_blocks_total gauge", f"cb200_kv_blocks_total {cm.total_num_blocks}",
                 "# TYPE cb200_kv_blocks_free gauge", f"cb200_kv_blocks_free {cm.num_available_blocks}"]
        return Response(content="\n".join(lines) + "\n", media_type="text/plain; version=0.0.4")

    @app.get("/engine_check")
    def engine_check():
        return JSONResponse({"status": "Running" if async_engine.background_loop_status else "Error"})

    @app.post("/generate")
    async def generate(request: Request):
        body = await request.json()
        prompt = body.pop("prompt")
        stream = str(body.pop("stream", "false")).lower()
        request_id = id_generator()
        results = async_engine.generate(request_id, prompt, generation_config=get_generation_config(body))

        async def stream_results():
            async for out in results:
                yield (json.dumps({"text": out}) + "\0").encode("utf-8")

        if stream == "true":
            return StreamingResponse(stream_results())
        final = None
        async for out in results:
            if await request.is_disconnected():
                await async_engine.abort(request_id)
                return Response(status_code=499)
            final = out
        return JSONResponse({"text": final})

    @app.post("/completion")
    async def create_completion(request: Request):
        body = await request.json()
        stream = str(body.get("stream", "false")).lower()
        gen = get_generation_config({k: v for k, v in body.items() if k not in ("prompt", "stream")})
        result = await completion_serving.create_completion(request, gen)
        if stream == "true":
            return StreamingResponse(content=iter([json.dumps(result) + "\n\n"]), media_type="text/event-stream")
        return JSONResponse(content=result)

    @app.post("/chat")
    async def create_chat(request: Request):
        body = await request.json()
        stream = str(body.get("stream", "false")).lower()
        gen = get_generation_config({k: v for k, v in body.items() if k not in ("messages", "stream")})
        message = await chat_serving.create_chat(request, gen)
        if stream == "true":
            return StreamingResponse(content=message, media_type="text/event-stream")
        return JSONResponse(content=message)

    return app


def add_engine_config(parser: argparse.ArgumentParser) -> argparse.ArgumentParser:
    parser.add_argument("-m", "--model", type=str, default="llama-tiny",
                        help="model zoo name or path of a HF-format checkpoint directory")
    parser.add_argument("--block_size", type=int, default=16, choices=[16, 32, 64], help="token block size")
    parser.add_argument("--max_batch_size", type=int, default=8)
    parser.add_argument("-i", "--max_input_len", type=int, default=128)
    parser.add_argument("-o", "--max_output_len", type=int, default=128)
    parser.add_argument("-d", "--dtype", type=str, default="bf16", choices=["fp16", "fp32", "bf16"])
    parser.add_argument("--use_cuda_graph", action="store_true")
    parser.add_argument("--prompt_template", choices=prompt_template_choices, default=None)
    return parser


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="colossalai_b200 inference HTTP server")
    parser.add_argument("--host", type=str, default="127.0.0.1")
    parser.add_argument("--port", type=int, default=8000)
    parser.add_argument("--ssl-keyfile", type=str, default=None)
    parser.add_argument("--ssl-certfile", type=str, default=None)
    parser.add_argument("--root-path", type=str, default=None)
    parser.add_argument("--model_name", type=str, default=None)
    parser.add_argument("--chat-template", type=str, default=None)
    parser.add_argument("--response-role", type=str, default="assistant")
    return add_engine_config(parser).parse_args(argv)


def main(argv=None) -> None:
    import os

    import uvicorn

    from ...models import MODEL_ZOO, build_model

    args = parse_args(argv)
    cfg = InferenceConfig(max_batch_size=args.max_batch_size, max_input_len=args.max_input_len,
                          max_output_len=args.max_output_len, block_size=args.block_size, dtype=args.dtype,
                          use_cuda_graph=args.use_cuda_graph, prompt_template=args.prompt_template)
    tokenizer = None
    if os.path.isdir(args.model):
        model = args.model
        try:
            from transformers import AutoTokenizer

            tokenizer = AutoTokenizer.from_pretrained(args.model)
        except Exception:
            tokenizer = None
    else:
        assert args.model in MODEL_ZOO, f"unknown model {args.model}"
        model = build_model(args.model)
    engine = AsyncInferenceEngine(start_engine_loop=True, model_or_path=model, tokenizer=tokenizer,
                                  inference_config=cfg)
    app = build_app(engine, args.model_name or args.model, tokenizer, args.response_role, args.chat_template)
    app.root_path = args.root_path or ""
    uvicorn.run(app=app, host=args.host, port=args.port, log_level="info", timeout_keep_alive=TIMEOUT_KEEP_ALIVE,
                ssl_keyfile=args.ssl_keyfile, ssl_certfile=args.ssl_certfile)


if __name__ == "__main__":
    main()
