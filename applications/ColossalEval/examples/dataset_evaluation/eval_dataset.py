"""Phase 2 of a dataset evaluation: score the saved answers (no model, no GPU) and print / save the table.
Parity: reference `applications/ColossalEval/examples/dataset_evaluation/eval_dataset.py`."""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", "..", ".."))

from colossal_eval import format_table, run_evaluation  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inference_dir", required=True)
    ap.add_argument("--config", default=None, help="the inference config (its `metrics` section selects text metrics)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    metrics = json.load(open(args.config)).get("metrics") if args.config else None
    results = run_evaluation(args.inference_dir, metrics, args.out)
    print(format_table(results))


if __name__ == "__main__":
    main()
