"""Micro-benchmarks of the inference kernels against their plain PyTorch formulations: RMSNorm (+ residual), SwiGLU,
RoPE on the packed QKV, paged KV-cache write, split-KV paged decode attention, fp8 KV conversion.

    python examples/inference/benchmark_ops/benchmark_ops.py                       # all ops, Llama-3-8B shapes
    python examples/inference/benchmark_ops/benchmark_ops.py --ops decode rmsnorm --batch 64 --context 4096

Every row: device time of the native kernel (CUDA events, after warm-up, an L2-sized buffer is rewritten between
timed calls), the same for the PyTorch formulation, the speed-up, the bytes the op has to move and what fraction of
the HBM bandwidth the native kernel reaches (the ops are all memory-bound; the denominator is `--hbm_gbps`, default the
copy bandwidth `hbm_gbs` of `MEASURED_PEAKS.json` if present).  On a machine without a GPU the script only checks that the two
formulations agree on small shapes (the native wrappers fall back to PyTorch there) and says so.

Parity: reference `examples/inference/benchmark_ops/benchmark_{rmsnorm,rotary_embedding,decoding_attn,kv_cache_memcopy,
context_attn_unpad,fused_rotary_embdding_unpad,xine_copy}.py` (triton.testing perf reports of Triton vs CUDA vs torch).
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))

from colossalai_b200.ops import inference as infer_ops  # noqa: E402
from colossalai_b200.ops.activation import glu, glu_ref  # noqa: E402
from colossalai_b200.ops.norm import rms_norm, rms_norm_ref  # noqa: E402
from colossalai_b200.ops.rope import build_rope_cache, rope_qkv, rope_ref  # noqa: E402

CUDA = torch.cuda.is_available()
DEV = torch.device("cuda", 0) if CUDA else torch.device("cpu")


def timed(fn, iters: int = 20, warmup: int = 5) -> float:
    """Mean device milliseconds of `fn()`; L2 is flushed between timed calls."""
    if not CUDA:
        fn()
        return float("nan")
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / iters


def paged_cache(batch, context, bs, Hkv, D, dtype):
    blocks_per_seq = (context + bs - 1) // bs
    nb = batch * blocks_per_seq
    k = torch.randn(nb, bs, Hkv, D, device=DEV, dtype=dtype)
    v = torch.randn(nb, bs, Hkv, D, device=DEV, dtype=dtype)
    tables = torch.randperm(nb, device=DEV).view(batch, blocks_per_seq).int()
    return k, v, tables


def bench_rmsnorm(a):
    x = torch.randn(a.tokens, a.hidden, device=DEV, dtype=a.dtype)
    r = torch.randn_like(x)
    w = torch.randn(a.hidden, device=DEV, dtype=a.dtype)
    rows = []
    for name, res in (("rmsnorm", None), ("add+rmsnorm", r)):
        nat = lambda: rms_norm(x, w, 1e-6, res)
        ref = lambda: rms_norm_ref(x, w, 1e-6, res)
        o1, o2 = nat(), ref()
        o1, o2 = (o1[0], o2[0]) if res is not None else (o1, o2)
        torch.testing.assert_close(o1.float(), o2.float(), atol=3e-2, rtol=3e-2)
        n = x.numel() * x.element_size()
        rows.append((name, timed(nat), timed(ref), n * (2 if res is None else 4)))
    return rows


def bench_swiglu(a):
    gu = torch.randn(a.tokens, 2 * a.intermediate, device=DEV, dtype=a.dtype)
    torch.testing.assert_close(glu(gu).float(), glu_ref(gu).float(), atol=3e-2, rtol=3e-2)
    return [("silu_and_mul", timed(lambda: glu(gu)), timed(lambda: glu_ref(gu)), gu.numel() * gu.element_size() * 3 // 2)]


def bench_rope(a):
    width = (a.heads + 2 * a.kv_heads) * a.head_dim
    qkv = torch.randn(a.tokens, width, device=DEV, dtype=a.dtype)
    cos, sin = build_rope_cache(max(a.context, a.tokens) + 1, a.head_dim, device=DEV)
    pos = torch.randint(0, a.context, (a.tokens,), device=DEV)

    def ref():
        q, k, v = qkv.split([a.heads * a.head_dim, a.kv_heads * a.head_dim, a.kv_heads * a.head_dim], -1)
        q = rope_ref(q.reshape(a.tokens, a.heads, a.head_dim), pos, cos, sin)
        k = rope_ref(k.reshape(a.tokens, a.kv_heads, a.head_dim), pos, cos, sin)
        return torch.cat([q.reshape(a.tokens, -1), k.reshape(a.tokens, -1), v], -1)

    nat = lambda: rope_qkv(qkv.clone(), pos, cos, sin, a.heads, a.kv_heads, a.head_dim)
    torch.testing.assert_close(nat().float(), ref().float(), atol=3e-2, rtol=3e-2)
    moved = a.tokens * (a.heads + a.kv_heads) * a.head_dim * qkv.element_size() * 2
    return [("rope(q,k) on packed qkv", timed(nat), timed(ref), moved)]


def bench_kv_write(a):
    k_cache, v_cache, tables = paged_cache(a.batch, a.context, a.block_size, a.kv_heads, a.head_dim, a.dtype)
    tokens = a.batch                                      # decode: one new token per sequence
    k = torch.randn(tokens, a.kv_heads, a.head_dim, device=DEV, dtype=a.dtype)
    v = torch.randn_like(k)
    seq = torch.arange(a.batch, device=DEV, dtype=torch.int32)
    pos = torch.full((a.batch,), a.context - 1, device=DEV, dtype=torch.int32)

    def ref():
        blk = tables[seq.long(), (pos // a.block_size).long()].long()
        k_cache[blk, (pos % a.block_size).long()] = k
        v_cache[blk, (pos % a.block_size).long()] = v

    nat = lambda: infer_ops.kv_cache_write(k, v, k_cache, v_cache, tables, seq, pos)
    nat()
    blk = tables[seq.long(), (pos // a.block_size).long()].long()
    torch.testing.assert_close(k_cache[blk, (pos % a.block_size).long()], k)
    rows = [("kv write (decode step)", timed(nat), timed(ref), 4 * k.numel() * k.element_size())]
    # prefill: every prompt token of every sequence
    T = a.batch * min(a.context, 1024)
    kp = torch.randn(T, a.kv_heads, a.head_dim, device=DEV, dtype=a.dtype)
    vp = torch.randn_like(kp)
    seqp = torch.arange(a.batch, device=DEV, dtype=torch.int32).repeat_interleave(T // a.batch)
    posp = torch.arange(T // a.batch, device=DEV, dtype=torch.int32).repeat(a.batch)
    natp = lambda: infer_ops.kv_cache_write(kp, vp, k_cache, v_cache, tables, seqp, posp)

    def refp():
        b = tables[seqp.long(), (posp // a.block_size).long()].long()
        k_cache[b, (posp % a.block_size).long()] = kp
        v_cache[b, (posp % a.block_size).long()] = vp

    rows.append((f"kv write (prefill, {T} tokens)", timed(natp), timed(refp), 4 * kp.numel() * kp.element_size()))
    return rows


def bench_decode(a):
    k_cache, v_cache, tables = paged_cache(a.batch, a.context, a.block_size, a.kv_heads, a.head_dim, a.dtype)
    q = torch.randn(a.batch, a.heads, a.head_dim, device=DEV, dtype=a.dtype)
    lens = torch.full((a.batch,), a.context, device=DEV, dtype=torch.int32)
    nat = lambda: infer_ops.paged_decode_attention(q, k_cache, v_cache, tables, lens)
    ref = lambda: infer_ops.paged_decode_attention_ref(q, k_cache, v_cache, tables, lens)
    torch.testing.assert_close(nat().float(), ref().float(), atol=3e-2, rtol=3e-2)
    moved = 2 * a.batch * a.context * a.kv_heads * a.head_dim * q.element_size()
    return [(f"paged decode attention (b{a.batch}, ctx {a.context}, GQA {a.heads}/{a.kv_heads})", timed(nat),
             timed(ref, iters=3, warmup=1), moved)]


def bench_fp8(a):
    x = torch.randn(a.batch * 1024, a.kv_heads * a.head_dim, device=DEV, dtype=a.dtype)
    nat = lambda: infer_ops.convert_fp8(x, True)
    ref = lambda: x.to(torch.float8_e5m2)
    back = infer_ops.convert_fp8(nat(), False, a.dtype)
    torch.testing.assert_close(back.float(), x.float(), atol=0.3, rtol=0.3)
    return [("kv -> fp8 (e5m2)", timed(nat), timed(ref), x.numel() * (x.element_size() + 1))]


OPS = {"rmsnorm": bench_rmsnorm, "swiglu": bench_swiglu, "rope": bench_rope, "kv_write": bench_kv_write,
       "decode": bench_decode, "fp8": bench_fp8}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", nargs="*", default=list(OPS), choices=list(OPS))
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--context", type=int, default=2048)
    ap.add_argument("--tokens", type=int, default=8192, help="rows of the token-major ops (norm, swiglu, rope)")
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--intermediate", type=int, default=14336)
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--kv_heads", type=int, default=8)
    ap.add_argument("--head_dim", type=int, default=128)
    ap.add_argument("--block_size", type=int, default=16)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--hbm_gbps", type=float, default=None)
    a = ap.parse_args()
    a.dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float16
    if not CUDA:                                         # correctness-only tier: small shapes, fp32-friendly tolerances
        a.batch, a.context, a.tokens, a.hidden, a.intermediate = 2, 64, 16, 64, 128
        a.heads, a.kv_heads, a.head_dim = 4, 2, 16
        print("no GPU: checking native wrappers against the PyTorch formulations on small shapes (no timings)")
    bw = a.hbm_gbps
    peaks = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "MEASURED_PEAKS.json")
    if bw is None and os.path.exists(peaks):
        try:
            d = json.load(open(peaks))
            bw = float(d.get("hbm_gbs") or 0) or None
        except Exception:
            bw = None
    bw = bw or 6560.0
    print(f"{'op':58s} {'native ms':>10s} {'torch ms':>10s} {'speed-up':>9s} {'GB moved':>9s} {'% of HBM':>9s}")
    for name in a.ops:
        for label, t_nat, t_ref, nbytes in OPS[name](a):
            if math.isnan(t_nat):
                print(f"{label:58s} {'ok':>10s}")
                continue
            print(f"{label:58s} {t_nat:10.4f} {t_ref:10.4f} {t_ref / t_nat:8.2f}x {nbytes / 1e9:9.3f} "
                  f"{100.0 * nbytes / (t_nat * 1e-3) / (bw * 1e9):8.1f}%")


if __name__ == "__main__":
    main()
