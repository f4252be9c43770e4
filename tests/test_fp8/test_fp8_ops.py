"""FP8 cast / collectives / linear / hook + weight-only quantisation (reference: tests/test_fp8/*.py)."""
import pytest
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

import colossalai_b200
from colossalai_b200.quantization import (BnbQuantizationConfig, FP8Hook, all_gather_fp8, all_reduce_fp8,
                                          all_to_all_fp8, all_to_all_single_fp8, cast_from_fp8, cast_to_fp8, linear_fp8,
                                          quantize_model, reduce_scatter_fp8)
from colossalai_b200.tensor import ColoParameter, ColoParamOpHookManager
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


@pytest.mark.parametrize("fmt", ["e4m3", "e5m2"])
@pytest.mark.parametrize("per_channel", [False, True])
def test_cast_roundtrip_cpu(fmt, per_channel):
    torch.manual_seed(0)
    x = torch.randn(64, 96) * 3
    q, sinv = cast_to_fp8(x, fmt, per_channel_scale=per_channel)
    back = cast_from_fp8(q, sinv, torch.float32, per_channel_scale=per_channel)
    tol = 0.07 if fmt == "e4m3" else 0.13
    assert (back - x).abs().max() <= tol * x.abs().max()


def test_linear_fp8_and_hook_cpu():
    torch.manual_seed(0)
    x = torch.randn(32, 64, dtype=torch.bfloat16, requires_grad=True)
    w = torch.randn(48, 64, dtype=torch.bfloat16, requires_grad=True)
    y = linear_fp8(x, w)
    ref = F.linear(x.float(), w.float())
    assert (y.float() - ref).abs().max() < 0.08 * ref.abs().max()
    y.float().sum().backward()
    assert x.grad.shape == x.shape and w.grad.shape == w.shape
    gw_ref = torch.ones(32, 48).t() @ x.detach().float()
    assert (w.grad.float() - gw_ref).abs().max() < 0.1 * gw_ref.abs().max()
    # hook rewrites F.linear on ColoParameters
    p = ColoParameter(w.detach().clone())
    calls = []
    import colossalai_b200.quantization.fp8_hook as fh

    orig = fh.linear_fp8
    fh.linear_fp8 = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        with ColoParamOpHookManager.use_hooks(FP8Hook()):
            out = F.linear(x.detach(), p)
    finally:
        fh.linear_fp8 = orig
    assert calls and out.shape == (32, 48)


@pytest.mark.parametrize("mode", ["int8", "nf4", "fp4"])
def test_weight_only_quantisation(mode):
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 8))
    x = torch.randn(4, 64)
    ref = model(x)
    cfg = (BnbQuantizationConfig(load_in_8bit=True, torch_dtype=torch.float32) if mode == "int8" else
           BnbQuantizationConfig(load_in_4bit=True, bnb_4bit_quant_type=mode, bnb_4bit_compute_dtype="fp32",
                                 bnb_4bit_use_double_quant=(mode == "nf4")))
    qm = quantize_model(model, cfg).cpu()
    assert isinstance(qm[4], nn.Linear) and not isinstance(qm[0], nn.Linear)     # head kept, body quantised
    out = qm(x)
    tol = 0.05 if mode == "int8" else 0.5
    assert (out - ref).abs().max() < tol * ref.abs().max().clamp_min(1.0)


def _dist_worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    torch.manual_seed(rank)
    x = torch.randn(1000) * 2
    ref = x.clone()
    dist.all_reduce(ref)
    all_reduce_fp8(x, "e4m3")
    assert (x - ref).abs().max() < 0.15 * ref.abs().max()
    # all_gather
    t = torch.randn(33, 7)
    outs = [torch.empty_like(t) for _ in range(world_size)]
    refs = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(refs, t)
    all_gather_fp8(outs, t, fp8_format="e4m3")
    for o, r in zip(outs, refs):
        assert (o - r).abs().max() < 0.07 * r.abs().max()
    # reduce_scatter
    ins = [torch.randn(16, 5) for _ in range(world_size)]
    out, ref = torch.empty(16, 5), torch.empty(16, 5)
    dist.reduce_scatter(ref, [i.clone() for i in ins]) if dist.get_backend() != "gloo" else None
    full = torch.stack(ins)
    dist.all_reduce(full)
    ref = full[rank]
    reduce_scatter_fp8(out, ins, fp8_format="e4m3")
    assert (out - ref).abs().max() < 0.15 * ref.abs().max()
    # all_to_all_single (even) + list variant (uneven)
    src = torch.randn(world_size * 4, 6)
    dst, dref = torch.empty_like(src), torch.empty_like(src)
    dist.all_to_all_single(dref, src)
    all_to_all_single_fp8(dst, src, fp8_format="e4m3")
    assert (dst - dref).abs().max() < 0.07 * src.abs().max() * 1.5
    ins = [torch.full((rank + 1 + r,), float(rank * 10 + r)) for r in range(world_size)]
    outs = [torch.empty(r + 1 + rank) for r in range(world_size)]
    all_to_all_fp8(outs, ins, fp8_format="e5m2")
    for r in range(world_size):
        assert torch.allclose(outs[r], torch.full((r + 1 + rank,), float(r * 10 + rank)), rtol=0.13)
    dist.barrier()
    dist.destroy_process_group()


@rerun_if_address_is_in_use()
def test_fp8_collectives_gloo():
    spawn(_dist_worker, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["e4m3", "e5m2"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
def test_native_fp8_cast_matches_reference(fmt, dtype):
    from colossalai_b200.ops._dispatch import force_torch

    torch.manual_seed(0)
    x = (torch.randn(257, 1000, device="cuda") * 4).to(dtype)
    for pc in (False, True):
        q, s = cast_to_fp8(x, fmt, per_channel_scale=pc)
        with force_torch():
            q_ref, s_ref = cast_to_fp8(x, fmt, per_channel_scale=pc)
        torch.testing.assert_close(s.reshape(-1), s_ref.reshape(-1).float(), rtol=1e-6, atol=0)
        mism = (q.view(torch.uint8) != q_ref.view(torch.uint8)).float().mean().item()
        assert mism < 2e-3, mism        # fast-math rounding may flip a handful of ties
        back = cast_from_fp8(q, s, dtype, per_channel_scale=pc)
        with force_torch():
            back_ref = cast_from_fp8(q, s, dtype, per_channel_scale=pc)
        torch.testing.assert_close(back.float(), back_ref.float(), rtol=1e-2, atol=1e-3)


@pytest.mark.gpu
def test_linear_fp8_gpu():
    torch.manual_seed(0)
    x = torch.randn(256, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(384, 512, device="cuda", dtype=torch.bfloat16) * 0.05).requires_grad_()
    y = linear_fp8(x, w)
    ref = F.linear(x.float(), w.float())
    assert (y.float() - ref).abs().max() < 0.1 * ref.abs().max()
    y.float().square().mean().backward()
    gy = (2 * ref / ref.numel())
    gx_ref = gy @ w.float()
    assert (x.grad.float() - gx_ref).abs().max() < 0.15 * gx_ref.abs().max()
