"""Tiny drivers for `ncu` captures: each target launches ONE kernel family a few times on representative shapes.
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1 -o gpurun_out/ncu_<t> python scripts/ncu_targets.py <t>
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def gemm():
    from colossalai_b200.ops import gemm_native as g
    x = torch.randn(8192, 4096, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(14336, 4096, device="cuda", dtype=torch.bfloat16)
    for _ in range(4):
        g.gemm_nt(x, w)


def flash_fwd():
    from colossalai_b200.ops import flash_attn_native as fa
    q, k, v = (torch.randn(4096, h, 128, device="cuda", dtype=torch.bfloat16) for h in (32, 8, 8))
    for _ in range(4):
        fa.flash_fwd(q, k, v, 1, True, None)


def flash_bwd():
    from colossalai_b200.ops import flash_attn_native as fa
    q, k, v = (torch.randn(4096, h, 128, device="cuda", dtype=torch.bfloat16) for h in (32, 8, 8))
    o, lse = fa.flash_fwd(q, k, v, 1, True, None)
    do = torch.randn_like(o)
    for _ in range(4):
        fa.flash_bwd(q, k, v, o, do, lse, 1, True, None)


def grouped():
    from colossalai_b200.moe.grouped_gemm import grouped_linear
    c = torch.tensor([2048] * 8, device="cuda")
    x = torch.randn(16384, 4096, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(8, 14336, 4096, device="cuda", dtype=torch.bfloat16) * 0.02
    for _ in range(4):
        grouped_linear(x, w, c)


def norm_glu():
    from colossalai_b200 import ops
    x = torch.randn(8192, 4096, device="cuda", dtype=torch.bfloat16)
    w = torch.ones(4096, device="cuda", dtype=torch.bfloat16)
    h = torch.randn(8192, 28672, device="cuda", dtype=torch.bfloat16)
    for _ in range(4):
        ops.rms_norm(x, w, 1e-5)
        ops.glu(h, "silu")


def adam():
    from colossalai_b200.nn.optimizer import FusedAdam
    p = [torch.nn.Parameter(torch.randn(64 << 20, device="cuda")) for _ in range(4)]
    for q in p:
        q.grad = torch.randn_like(q)
    opt = FusedAdam(p, lr=1e-3)
    for _ in range(4):
        opt.step()


def decode():
    from colossalai_b200.ops import inference as iops
    bsz, Hq, Hkv, D, bs, ctx = 64, 32, 8, 128, 64, 2048
    nb = bsz * ctx // bs
    kc = torch.randn(nb, bs, Hkv, D, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn_like(kc)
    bt = torch.arange(nb, device="cuda", dtype=torch.int32).view(bsz, -1)
    sl = torch.full((bsz,), ctx, device="cuda", dtype=torch.int32)
    q = torch.randn(bsz, Hq, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(4):
        iops.paged_decode_attention(q, kc, vc, bt, sl, 0.088)


def moe():
    from colossalai_b200.moe import dispatch_combine as dc
    from colossalai_b200.moe.grouped_gemm import grouped_linear
    dc.set_moe_backend("fused")
    x = torch.randn(8192, 4096, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(8, 4096, 4096, device="cuda", dtype=torch.bfloat16) * 0.02
    lg = torch.randn(8192, 8, device="cuda")
    for _ in range(4):
        tw, ti = lg.softmax(-1).topk(2, -1)
        dc.moe_forward(x, tw, ti, lambda r, c: grouped_linear(r, w, c), 8, None)


def ce_rope():
    from colossalai_b200.ops import cross_entropy as ce
    logits = torch.randn(8192, 128256, device="cuda", dtype=torch.bfloat16)
    tgt = torch.randint(0, 128256, (8192,), device="cuda")
    one = torch.ones((), device="cuda")
    for _ in range(4):
        gmax = ce.row_max(logits)
        st = ce.sumexp_and_target(logits, tgt, gmax, 0, -100)
        ce.softmax_grad(logits, tgt, gmax, st[0], one, 0, -100)


def kv_write():
    from colossalai_b200.ops import inference as iops
    T, Hkv, D, bs = 16384, 8, 128, 64
    k = torch.randn(T, Hkv, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn_like(k)
    kc = torch.zeros(T // bs, bs, Hkv, D, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    bt = torch.arange(T // bs, device="cuda", dtype=torch.int32).view(16, -1)
    seq = torch.arange(T, device="cuda", dtype=torch.int32) // (T // 16)
    pos = torch.arange(T, device="cuda", dtype=torch.int32) % (T // 16)
    for _ in range(4):
        iops.kv_cache_write(k, v, kc, vc, bt, seq, pos)


if __name__ == "__main__":
    {"gemm": gemm, "flash_fwd": flash_fwd, "flash_bwd": flash_bwd, "grouped": grouped, "norm_glu": norm_glu, "adam": adam,
     "decode": decode, "moe": moe, "ce_rope": ce_rope, "kv_write": kv_write}[sys.argv[1]]()
    torch.cuda.synchronize()
