"""Offline generation with the paged-KV continuous-batching engine (reference: examples/inference/llama/llama_generation.py).

    python examples/inference/llama_generation.py -m llama-tiny --max_length 32 [--spec] [--cuda_graph]
    python examples/inference/llama_generation.py -m /path/to/hf_llama_dir -p "Hello" --dtype bf16
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

from colossalai_b200.inference import GenerationConfig, InferenceConfig, InferenceEngine  # noqa: E402
from colossalai_b200.models import MODEL_ZOO, build_model, get_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-m", "--model", default="llama-tiny", help="model zoo name or HF checkpoint directory")
    ap.add_argument("-p", "--prompt", action="append", default=None)
    ap.add_argument("--max_length", type=int, default=32, help="new tokens per request")
    ap.add_argument("-b", "--max_batch_size", type=int, default=8)
    ap.add_argument("--dtype", default="bf16" if torch.cuda.is_available() else "fp32")
    ap.add_argument("--cuda_graph", action="store_true")
    ap.add_argument("--spec", action="store_true", help="speculative decoding with a 1-layer drafter of the same family")
    ap.add_argument("--do_sample", action="store_true")
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--top_k", type=int, default=None)
    ap.add_argument("--top_p", type=float, default=None)
    args = ap.parse_args()
    prompts = args.prompt or ["Introduce some landmarks in Beijing", "The capital of France is", "def fibonacci(n):"]
    tokenizer = None
    if os.path.isdir(args.model):
        model = args.model
        try:
            from transformers import AutoTokenizer

            tokenizer = AutoTokenizer.from_pretrained(args.model)
        except Exception:
            pass
    else:
        assert args.model in MODEL_ZOO, f"unknown model {args.model}"
        model = build_model(args.model)
    cfg = InferenceConfig(max_batch_size=args.max_batch_size, max_input_len=256, max_output_len=args.max_length,
                          dtype=args.dtype, use_cuda_graph=args.cuda_graph, block_size=16)
    engine = InferenceEngine(model, tokenizer, cfg, verbose=True)
    if args.spec:
        base = engine.model_config
        drafter = build_model(get_config(args.model, num_hidden_layers=1) if args.model in MODEL_ZOO else base)
        engine.enable_spec_dec(drafter, n_spec_tokens=4)
    gen = GenerationConfig(max_new_tokens=args.max_length, do_sample=args.do_sample, temperature=args.temperature,
                           top_k=args.top_k, top_p=args.top_p)
    t0 = time.perf_counter()
    outs = engine.generate(prompts=prompts, generation_config=gen)
    dt = time.perf_counter() - t0
    for p, o in zip(prompts, outs):
        print(f"--- {p!r}\n{o!r}")
    print(f"{len(prompts)} requests, {args.max_length} new tokens each, {dt:.2f} s")


if __name__ == "__main__":
    main()
