"""Serve a model over HTTP and talk to it: health check, plain generation, OpenAI-style completion and chat, the
Prometheus metrics page, and a small concurrent load (continuous batching at work).

    # terminal 1: the server (any zoo name or a HuggingFace checkpoint directory)
    python -m colossalai_b200.inference.server.api_server --model llama-tiny --port 8000 --dtype fp32
    # terminal 2: this client
    python examples/inference/client/run_client.py --url http://127.0.0.1:8000

    # or both in one process (starts the server on a free port in a background thread, then runs the client against it)
    python examples/inference/client/run_client.py --self-host --model llama-tiny

Parity: reference `examples/inference/client/{run_locust.sh, locustfile.py, test_ci.sh}` (locust load over /generate,
/completion, /chat of `colossalai.inference.server.api_server`).
"""
import argparse
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import requests

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))


def self_host(model_name: str, dtype: str):
    """The server's own `build_app` on a uvicorn thread; returns (base url, stop function)."""
    import torch
    import uvicorn

    from colossalai_b200.inference.config import InferenceConfig
    from colossalai_b200.inference.core.async_engine import AsyncInferenceEngine
    from colossalai_b200.inference.server.api_server import build_app
    from colossalai_b200.models import build_model
    from colossalai_b200.testing import free_port

    torch.manual_seed(0)
    model = build_model(model_name).eval()
    model = model.float() if dtype == "fp32" else model
    cfg = InferenceConfig(max_batch_size=8, max_input_len=64, max_output_len=32, block_size=16, dtype=dtype)
    engine = AsyncInferenceEngine(start_engine_loop=True, model_or_path=model, tokenizer=None, inference_config=cfg)
    port = free_port()
    server = uvicorn.Server(uvicorn.Config(build_app(engine, model_name), host="127.0.0.1", port=port, log_level="warning"))
    th = threading.Thread(target=server.run, daemon=True)
    th.start()
    url = f"http://127.0.0.1:{port}"
    for _ in range(100):
        try:
            if requests.get(url + "/ping", timeout=1).status_code == 200:
                break
        except requests.RequestException:
            time.sleep(0.1)

    def stop():
        server.should_exit = True
        th.join(timeout=10)

    return url, stop


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--url", default="http://127.0.0.1:8000")
    ap.add_argument("--self-host", action="store_true")
    ap.add_argument("--model", default="llama-tiny")
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--concurrency", type=int, default=8)
    ap.add_argument("--max_new_tokens", type=int, default=8)
    args = ap.parse_args()
    stop = None
    url = args.url
    if args.self_host:
        url, stop = self_host(args.model, args.dtype)
    try:
        print("ping      ", requests.get(url + "/ping").json())
        r = requests.post(url + "/generate", json={"prompt": "hello world", "max_new_tokens": args.max_new_tokens})
        print("generate  ", json.dumps(r.json())[:120])
        r = requests.post(url + "/completion", json={"prompt": "once upon a time", "max_new_tokens": args.max_new_tokens})
        print("completion", json.dumps(r.json())[:160])
        r = requests.post(url + "/chat", json={"messages": [{"role": "user", "content": "hi"}],
                                               "max_new_tokens": args.max_new_tokens})
        print("chat      ", json.dumps(r.json()["choices"][0]["message"])[:120])

        def one(i):
            t0 = time.perf_counter()
            rr = requests.post(url + "/generate", json={"prompt": f"request {i}", "max_new_tokens": args.max_new_tokens})
            return rr.status_code, time.perf_counter() - t0

        t0 = time.perf_counter()
        with ThreadPoolExecutor(args.concurrency) as pool:
            res = list(pool.map(one, range(4 * args.concurrency)))
        wall = time.perf_counter() - t0
        ok = sum(code == 200 for code, _ in res)
        lat = sorted(t for _, t in res)
        print(f"load       {ok}/{len(res)} ok, {len(res) / wall:.1f} req/s, p50 {lat[len(lat) // 2] * 1e3:.0f} ms, "
              f"p95 {lat[int(len(lat) * 0.95) - 1] * 1e3:.0f} ms at concurrency {args.concurrency}")
        metrics = requests.get(url + "/metrics").text
        print("metrics   ", [l for l in metrics.splitlines() if l.startswith("cb200_") and " " in l][:4])
        assert ok == len(res)
    finally:
        if stop is not None:
            stop()


if __name__ == "__main__":
    main()
