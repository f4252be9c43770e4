"""A/B of the CTA-pair GEMM's tail-wave schedules in ONE process (interleaved, L2 flushed): whole tiles vs 256 x 128 halves vs
K ranges for the partial last wave, next to cuBLAS, on the GEMMs of the N=1 Llama-3-8B step (M = 4096 tokens)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colossalai_b200.ops import gemm_native as g  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def once(fn):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e)


def main():
    M = int(os.environ.get("CB200_GEMM_BENCH_M", "4096"))
    shapes = [(M, 6144, 4096, "qkv"), (M, 4096, 4096, "o_proj"), (M, 4096, 14336, "down"), (M, 14336, 4096, "down dgrad"),
              (M, 28672, 4096, "gate_up")]
    for Mm, N, K, name in shapes:
        x = torch.randn(Mm, K, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        dy = torch.randn(Mm, N, device="cuda", dtype=torch.bfloat16)
        fl = 2.0 * Mm * N * K
        cases = {"nt": (lambda: g.gemm_nt(x, w), lambda: torch.nn.functional.linear(x, w)),
                 "nn": (lambda: g.gemm_nn(dy, w), lambda: dy @ w),
                 "tn": (lambda: g.gemm_tn(dy, x), lambda: dy.t() @ x)}
        row = {"shape": [Mm, N, K], "name": name}
        for label, (ours, lib) in cases.items():
            ts = {"whole": [], "half": [], "ksplit2": [], "cublas": []}
            for rep in range(12):
                for key in ("whole", "half", "ksplit2"):
                    os.environ["CB200_GEMM_TAIL_HALF"] = "1" if key == "half" else "0"
                    os.environ["CB200_GEMM_TAIL_SPLIT"] = "2" if key == "ksplit2" else "0"
                    t = once(ours)
                    if rep >= 2:
                        ts[key].append(t)
                t = once(lib)
                if rep >= 2:
                    ts["cublas"].append(t)
            row[label] = {k: round(fl / sorted(v)[len(v) // 2] / 1e9) for k, v in ts.items()}
        print("GEMM_AB " + json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
