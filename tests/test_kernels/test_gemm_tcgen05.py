"""tcgen05 GEMM numerics vs fp32 torch.matmul for the three linear-layer GEMM flavours (NT / NN / TN), tails,
accumulate-into-C and both tile widths."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(*shape, dtype=torch.bfloat16):
    return (torch.randn(*shape, device="cuda") * 0.5).to(dtype)


SHAPES = [(128, 256, 64), (256, 512, 128), (1000, 768, 520), (4096, 4096, 4096), (336, 1024, 2048), (8192, 6144, 4096)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("block_n", [128, 256, 512])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gemm_nt(M, N, K, block_n, dtype):
    from colossalai_b200.ops import gemm_native as g

    M = 333 if M == 336 else M     # odd M is fine for K-major A

    torch.manual_seed(0)
    x, w = _mk(M, K, dtype=dtype), _mk(N, K, dtype=dtype)
    y = g.gemm_nt(x, w, block_n=block_n)
    ref = x.float() @ w.float().t()
    torch.testing.assert_close(y.float(), ref, atol=0.15 * (K / 1024) ** 0.5 + 0.05, rtol=2e-2)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("block_n", [128, 256, 512])
def test_gemm_nn(M, N, K, block_n):
    from colossalai_b200.ops import gemm_native as g

    torch.manual_seed(0)
    a, b = _mk(M, K), _mk(K, N)
    c = g.gemm_nn(a, b, block_n=block_n)
    ref = a.float() @ b.float()
    torch.testing.assert_close(c.float(), ref, atol=0.15 * (K / 1024) ** 0.5 + 0.05, rtol=2e-2)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("block_n", [128, 256, 512])
def test_gemm_tn(M, N, K, block_n):
    from colossalai_b200.ops import gemm_native as g

    torch.manual_seed(0)
    a, b = _mk(K, M), _mk(K, N)      # contraction over dim 0
    c = g.gemm_tn(a, b, block_n=block_n)
    ref = a.float().t() @ b.float()
    torch.testing.assert_close(c.float(), ref, atol=0.15 * (K / 1024) ** 0.5 + 0.05, rtol=2e-2)


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_tn_accumulate(out_dtype):
    from colossalai_b200.ops import gemm_native as g

    torch.manual_seed(0)
    a, b = _mk(2048, 512), _mk(2048, 768)
    acc = torch.randn(512, 768, device="cuda").to(out_dtype)
    ref = acc.float() + a.float().t() @ b.float()
    g.gemm_tn(a, b, out=acc, accumulate=True)
    torch.testing.assert_close(acc.float(), ref, atol=0.3, rtol=2e-2)


def test_linear_autograd_through_native_gemm():
    """The op-level entry points (`ops.linear_forward / matmul_nn / matmul_tn`) used by every parallel linear."""
    from colossalai_b200 import ops
    from colossalai_b200.kernel import loader
    from colossalai_b200.ops import gemm as gemm_ops
    from colossalai_b200.shardformer.layer._operation import linear_with_grad_accum

    torch.manual_seed(0)
    x = _mk(1024, 512).requires_grad_(True)
    w = _mk(768, 512).requires_grad_(True)
    old = gemm_ops.get_gemm_backend()
    gemm_ops.set_gemm_backend("native")
    before = loader.launch_counter.count
    try:
        y = linear_with_grad_accum(x, w)
        (y.float() * 0.01).sum().backward()
    finally:
        gemm_ops.set_gemm_backend(old)
    assert loader.launch_counter.count >= before + 3, "the tcgen05 GEMM must have run for fwd, dgrad and wgrad"
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    ((xr @ wr.t()) * 0.01).sum().backward()
    torch.testing.assert_close(y.float(), xr @ wr.t(), atol=0.2, rtol=2e-2)
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=5e-2, rtol=2e-2)
