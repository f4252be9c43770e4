"""tcgen05 flash-attention forward against an fp32 PyTorch reference (out and log-sum-exp), plus gradients.
First hardware run (round 2): 17/17 pass on a B200 (gpurun_out/c1_flash_test.log)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,S,Hq,Hkv,D", [(1, 128, 2, 2, 128), (2, 512, 8, 2, 128), (1, 1024, 4, 4, 64),
                                          (1, 4096, 32, 8, 128)])
@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_flash_fwd_matches_reference(B, S, Hq, Hkv, D, causal, dtype):
    from colossalai_b200.ops import flash_attn_native as fa
    from colossalai_b200.ops.attention import attention_with_lse_ref

    torch.manual_seed(0)
    q = torch.randn(B * S, Hq, D, device="cuda", dtype=dtype)
    k = torch.randn(B * S, Hkv, D, device="cuda", dtype=dtype)
    v = torch.randn(B * S, Hkv, D, device="cuda", dtype=dtype)
    out, lse = fa.flash_fwd(q, k, v, B, causal, None)
    ref_o, ref_lse = attention_with_lse_ref(q.float(), k.float(), v.float(), batch=B, causal=causal)
    torch.cuda.synchronize()
    torch.testing.assert_close(lse, ref_lse, atol=2e-3, rtol=2e-3)
    torch.testing.assert_close(out.float(), ref_o, atol=2e-2, rtol=2e-2)


def test_flash_native_gradients_match_library():
    from colossalai_b200.ops import flash_attn_native as fa

    torch.manual_seed(0)
    B, S, Hq, Hkv, D = 2, 512, 8, 2, 128
    q, k, v = (torch.randn(B * S, h, D, device="cuda", dtype=torch.bfloat16, requires_grad=True) for h in (Hq, Hkv, Hkv))
    fa.enable(True)
    try:
        out = fa.flash_attention(q, k, v, batch=B, causal=True)
        out.float().pow(2).mean().backward()
        got = [t.grad.float().clone() for t in (q, k, v)]
    finally:
        fa.enable(os.environ.get("CB200_FLASH_NATIVE", "0") == "1")
    for t in (q, k, v):
        t.grad = None
    from colossalai_b200.ops.attention import attention_ref

    attention_ref(q, k, v, B, True, None).float().pow(2).mean().backward()
    for g, t in zip(got, (q, k, v)):
        assert (g - t.grad.float()).abs().max() <= 3e-2 * t.grad.float().abs().max() + 1e-6


@pytest.mark.parametrize("B,S,Hq,Hkv", [(1, 128, 2, 2), (1, 256, 4, 1), (2, 512, 8, 2), (1, 2048, 8, 8), (1, 4096, 32, 8)])
@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_flash_bwd_matches_reference(B, S, Hq, Hkv, causal, dtype):
    """tcgen05 backward (dq, dk, dv) against autograd through the explicit fp32 softmax reference."""
    from colossalai_b200.ops import flash_attn_native as fa
    from colossalai_b200.ops.attention import attention_with_lse_ref

    D = 128
    torch.manual_seed(1)
    q = torch.randn(B * S, Hq, D, device="cuda", dtype=dtype)
    k = torch.randn(B * S, Hkv, D, device="cuda", dtype=dtype)
    v = torch.randn(B * S, Hkv, D, device="cuda", dtype=dtype)
    dout = torch.randn(B * S, Hq, D, device="cuda", dtype=dtype)
    out, lse = fa.flash_fwd(q, k, v, B, causal, None)
    dq, dk, dv = fa.flash_bwd(q, k, v, out, dout, lse, B, causal, None)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref_o, _ = attention_with_lse_ref(qf, kf, vf, batch=B, causal=causal)
    ref_o.backward(dout.float())
    torch.cuda.synchronize()
    for name, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        err = (got.float() - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err <= 2e-2 * scale + 1e-3, f"{name}: max err {err:.4g} vs max |ref| {scale:.4g}"
        # no systematic bias: the mean signed error is tiny compared with the mean magnitude
        assert (got.float() - ref).mean().abs().item() <= 2e-3 * ref.abs().mean().item() + 1e-5, name
