"""Phase 1 of a dataset evaluation: answer every item of every configured dataset and save the answers.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 applications/ColossalEval/examples/dataset_evaluation/inference.py \
        --config applications/ColossalEval/examples/dataset_evaluation/config.json --out_dir /tmp/eval/answers
    python applications/ColossalEval/examples/dataset_evaluation/eval_dataset.py --inference_dir /tmp/eval/answers \
        --config applications/ColossalEval/examples/dataset_evaluation/config.json --out /tmp/eval/results.json

The config names the model (a zoo preset or a HuggingFace checkpoint directory) and the datasets: `mmlu` (directory of
`<subject>_test.csv`), `gsm8k` / `cloze` / `loss` (jsonl), or `synthetic` (built in, for a smoke run without files).
Every rank answers its share of the items; rank 0 writes `<dataset>_inference.json`.
Parity: reference `applications/ColossalEval/examples/dataset_evaluation/inference.py`.
"""
import argparse
import glob
import json
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", "..", ".."))

import colossalai_b200  # noqa: E402
from colossal_eval import (EvalModel, cloze_items, few_shot_prompt, gsm8k_items, load_jsonl, load_mmlu_csv,  # noqa: E402
                           run_inference)
from colossalai_b200.models import build_model, get_config  # noqa: E402
from colossalai_b200.testing import free_port  # noqa: E402


def synthetic_items():
    mc = [{"instruction": f"Question: what is {a} + {b}?\nAnswer:", "choices": [f" {a + b}", f" {a + b + 1}", f" {a * b + 3}"],
           "answer": 0, "category": "addition" if a < 5 else "addition-large"} for a in range(2, 8) for b in (1, 3)]
    gen = [{"instruction": f"Repeat the word: cat{i}\n", "target": f"cat{i}", "category": "copy"} for i in range(6)]
    loss = [{"instruction": "The quick brown fox", "target": " jumps over the lazy dog", "calculate_loss": True,
             "category": "pangram"} for _ in range(4)]
    return {"synthetic_choice": mc, "synthetic_generation": gen, "synthetic_loss": loss}


def load_datasets(cfg):
    out = {}
    for d in cfg["datasets"]:
        kind, name = d["type"], d.get("name", d["type"])
        if kind == "synthetic":
            out.update(synthetic_items())
            continue
        if kind == "mmlu":
            items = [it for f in sorted(glob.glob(os.path.join(d["path"], "*_test.csv"))) for it in load_mmlu_csv(f)]
        elif kind == "gsm8k":
            items = gsm8k_items(load_jsonl(d["path"]))
        elif kind == "loss":
            items = [dict(r, calculate_loss=True) for r in load_jsonl(d["path"])]
        else:
            items = cloze_items(load_jsonl(d["path"]), category=name)
        shots = d.get("few_shot", 0)
        if shots:
            items = [few_shot_prompt(it, items[:shots]) for it in items[shots:]]
        out[name] = items[: d.get("limit")] if d.get("limit") else items
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--out_dir", required=True)
    args = ap.parse_args()
    if "RANK" in os.environ:
        colossalai_b200.launch_from_torch(backend="nccl" if torch.cuda.is_available() else "gloo")
    else:
        colossalai_b200.launch(0, 1, "127.0.0.1", free_port(), verbose=False)
    cfg = json.load(open(args.config))
    m = cfg["model"]
    torch.manual_seed(0)
    if os.path.isdir(m["name"]):
        from transformers import AutoTokenizer

        from colossalai_b200.models.hf_io import load_hf_checkpoint

        model = load_hf_checkpoint(m["name"])
        tok = AutoTokenizer.from_pretrained(m["name"])
        encode, decode = (lambda t: tok(t, add_special_tokens=False)["input_ids"]), (lambda ids: tok.decode(ids))
        eos, bos = tok.eos_token_id, tok.bos_token_id
    else:
        model = build_model(get_config(m["name"]))
        encode, eos, bos = (lambda t: [3 + b for b in t.encode()]), 2, 1
        decode = lambda ids: bytes(max(0, min(255, i - 3)) for i in ids if i >= 3).decode(errors="replace")   # noqa: E731
    dev = colossalai_b200.accelerator.get_accelerator().get_current_device()
    model = (model if torch.cuda.is_available() else model.float()).to(dev)
    em = EvalModel(model, encode, decode, batch_size=m.get("batch_size", 8), max_new_tokens=m.get("max_new_tokens", 32),
                   eos_token_id=eos, bos_token_id=bos)
    paths = run_inference(em, load_datasets(cfg), args.out_dir)
    if dist.get_rank() == 0:
        for name, p in paths.items():
            print(f"{name}: {p}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
