"""Scaled masked / causal softmax kernels vs fp32 PyTorch (reference: tests/test_kernels? (scaled softmax) +
nn/layer/scaled_softmax usage)."""
import pytest
import torch

from colossalai_b200.nn.layer import AttnMaskType, FusedScaleMaskSoftmax, MixedFusedLayerNorm
from colossalai_b200.ops.softmax import (scaled_causal_softmax, scaled_causal_softmax_ref, scaled_masked_softmax,
                                         scaled_masked_softmax_ref)


def test_module_cpu_fallback():
    torch.manual_seed(0)
    x = torch.randn(2, 3, 8, 8)
    m = FusedScaleMaskSoftmax(attn_mask_type=AttnMaskType.causal, scale=0.5)
    y = m(x, None)
    torch.testing.assert_close(y, scaled_causal_softmax_ref(x, 0.5))
    ln = MixedFusedLayerNorm(8)
    torch.testing.assert_close(ln(x), torch.nn.functional.layer_norm(x, (8,)), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("sq,sk", [(17, 17), (64, 128), (128, 1000), (33, 2500), (5, 4097)])
def test_scaled_softmax_gpu(dtype, sq, sk):
    torch.manual_seed(0)
    x = torch.randn(2, 3, sq, sk, device="cuda", dtype=dtype, requires_grad=True)
    mask = torch.rand(2, 1, sq, sk, device="cuda") < 0.2
    mask[..., 0] = False
    tol = dict(rtol=2e-2, atol=2e-3) if dtype != torch.float32 else dict(rtol=1e-4, atol=1e-6)
    for fn, ref in ((lambda t: scaled_masked_softmax(t, mask, 0.3), lambda t: scaled_masked_softmax_ref(t, mask, 0.3)),
                    (lambda t: scaled_masked_softmax(t, mask[:1], 0.3), lambda t: scaled_masked_softmax_ref(t, mask[:1], 0.3)),
                    (lambda t: scaled_causal_softmax(t, 0.7), lambda t: scaled_causal_softmax_ref(t, 0.7))):
        y = fn(x)
        xr = x.detach().float().requires_grad_()
        yr = ref(xr)
        torch.testing.assert_close(y.float(), yr, **tol)
        g = torch.randn_like(y)
        y.backward(g)
        yr.backward(g.float())
        torch.testing.assert_close(x.grad.float(), xr.grad, **tol)
        x.grad = None
