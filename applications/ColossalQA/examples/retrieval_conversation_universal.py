"""Bilingual retrieval conversation over your own files, answered by a local model.

    python applications/ColossalQA/examples/retrieval_conversation_universal.py --files notes.md faq_zh.txt products.csv \
        --model llama-tiny --ask "What is the warranty period?" "保修期是多久？"
    python applications/ColossalQA/examples/retrieval_conversation_universal.py          # built-in documents, interactive

Documents are split (Chinese on its own punctuation), indexed per language (BM25 + hashed embeddings fused by
reciprocal rank), and every question is answered from the retrieved chunks only; follow-ups use the conversation
memory.  The generator is the paged-KV `InferenceEngine` over a zoo preset / HF checkpoint directory (`--model`), or an
already running server (`--server http://127.0.0.1:8000`).  With the random-weight tiny preset the answers are noise -
the point of the smoke run is the retrieval and the plumbing; sources are printed with every answer.
Parity: reference `applications/ColossalQA/examples/retrieval_conversation_universal.py`.
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", ".."))

from colossalqa import (BM25Index, EmbeddingIndex, HybridRetriever, LocalLLM, UniversalRetrievalConversation,  # noqa: E402
                        load_table)

BUILT_IN = [
    "The Aurora X2 laptop has a warranty period of 24 months. Battery wear is covered for the first 12 months only.",
    "Returns are accepted within 30 days of delivery if the product is unused and in its original packaging.",
    "极光 X2 笔记本电脑的保修期为24个月。电池损耗仅在前12个月内保修。退货须在收货后30天内提出，且产品未使用、包装完好。",
]


def build_llm(args) -> LocalLLM:
    if args.server:
        return LocalLLM.from_http(args.server, max_new_tokens=args.max_new_tokens)
    import torch

    import colossalai_b200
    from colossalai_b200.inference import InferenceConfig, InferenceEngine
    from colossalai_b200.inference.config import GenerationConfig
    from colossalai_b200.models import build_model
    from colossalai_b200.testing import free_port

    colossalai_b200.launch(0, 1, "127.0.0.1", free_port(), verbose=False)
    torch.manual_seed(0)
    cuda = torch.cuda.is_available()
    model = build_model(args.model).eval()
    model = model if cuda else model.float()
    # without a tokenizer the engine works on UTF-8 bytes: keep the prompt's tail inside the model's context window
    limit = model.cfg.max_position_embeddings - args.max_new_tokens - 8
    engine = InferenceEngine(model, None, InferenceConfig(max_batch_size=1, max_input_len=limit, block_size=16,
                                                           max_output_len=args.max_new_tokens, dtype="bf16" if cuda else "fp32"))
    gen = GenerationConfig(max_new_tokens=args.max_new_tokens, do_sample=False)

    def tail(p: str) -> str:
        b = p.encode()[-(limit - 4):]
        return b.decode(errors="ignore")

    return LocalLLM(lambda p: engine.generate(prompts=[tail(p)], generation_config=gen))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", nargs="*", default=[])
    ap.add_argument("--model", default="llama-tiny")
    ap.add_argument("--server", default=None)
    ap.add_argument("--max_new_tokens", type=int, default=32)
    ap.add_argument("--ask", nargs="*", default=None)
    args = ap.parse_args()
    conv = UniversalRetrievalConversation(build_llm(args), k=2,
                                          make_index=lambda: HybridRetriever([BM25Index(), EmbeddingIndex()]))
    if args.files:
        for f in args.files:
            if f.endswith((".csv", ".jsonl", ".json")):
                print(f, conv.add_documents([d["text"] for d in load_table(f)], source=os.path.basename(f)))
            else:
                print(f, conv.add_documents([open(f).read()], source=os.path.basename(f)))
    else:
        print("built-in documents:", conv.add_documents(BUILT_IN, source="built-in"))
    questions = args.ask if args.ask is not None else iter(lambda: input("you> "), "")
    for q in questions:
        answer, sources, meta = conv.run(q)
        print(f"[{meta['language']}] {q}\n  answer : {answer!r}")
        for s in sources:
            print(f"  source : ({s['source']}) {s['text'][:80]}")


if __name__ == "__main__":
    main()
