"""tcgen05 flash-attention forward against an fp32 PyTorch reference (out and log-sum-exp), plus gradients through
the library backward.  Opt-in (`CB200_TEST_FLASH_NATIVE=1`) until the kernel has had its first run on a B200."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CB200_TEST_FLASH_NATIVE", "0") != "1",
                                 reason="kernel not yet validated on hardware: set CB200_TEST_FLASH_NATIVE=1")]


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
