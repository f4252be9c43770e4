"""MoE kernels + expert-parallel dispatch/combine (reference: tests/test_moe/test_kernel.py, test_moe_ep_tp.py,
tests/test_shardformer/test_model/test_shard_mixtral.py)."""
import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.moe import MoeCombine, MoeDispatch, moe_cumsum
from colossalai_b200.moe import dispatch_combine as dc
from colossalai_b200.ops import moe as mops
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _routing(s, e, c, device):
    torch.manual_seed(0)
    logits = torch.randn(s, e, device=device).softmax(-1)
    top1 = logits.argmax(-1)
    mask = torch.nn.functional.one_hot(top1, e).to(torch.int32)
    ranks = moe_cumsum(mask)
    mask = mask * (ranks < c)
    dest = (ranks * mask).sum(-1).to(torch.int32)
    return logits, mask, dest


def _legacy_roundtrip(device, dtype):
    s, e, c, h = 64, 4, 12, 32
    logits, mask, dest = _routing(s, e, c, device)
    tokens = torch.randn(s, h, device=device, dtype=dtype, requires_grad=True)
    lg = logits.clone().requires_grad_()
    x = MoeDispatch.apply(tokens, mask, dest, e * c)
    assert x.shape == (e, c, h)
    out = MoeCombine.apply(x * 2.0, lg, mask, dest, e * c)
    # dense reference
    kept = mask.sum(-1).bool()
    ref = torch.where(kept[:, None], 2.0 * tokens.float() * (logits * mask).sum(-1, keepdim=True), torch.zeros(()).to(device))
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
    out.float().sum().backward()
    gref = torch.where(kept[:, None], 2.0 * (logits * mask).sum(-1, keepdim=True).expand(s, h), torch.zeros(()).to(device))
    torch.testing.assert_close(tokens.grad.float(), gref, rtol=2e-2, atol=2e-2)
    assert lg.grad.shape == lg.shape


def test_legacy_dispatch_combine_cpu():
    _legacy_roundtrip("cpu", torch.float32)


def test_router_topk_cpu():
    torch.manual_seed(0)
    lg = torch.randn(50, 8)
    w, idx = mops.router_topk(lg, 2)
    p = lg.softmax(-1)
    rw, ridx = p.topk(2, -1)
    assert torch.equal(idx, ridx)
    torch.testing.assert_close(w, rw / rw.sum(-1, keepdim=True))


def _ep_worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    from colossalai_b200.models import build_model, get_config
    from colossalai_b200.models.moe import SparseMoE

    cfg = get_config("mixtral-tiny")
    torch.manual_seed(0)
    ref = SparseMoE(cfg).float()
    torch.manual_seed(0)
    ep = SparseMoE(cfg).float()

    class _SC:
        ep_group = dist.group.WORLD

    ep.setup_parallel(_SC())
    solo = [dist.new_group([r]) for r in range(world_size)]      # group(None) means WORLD: give the reference a 1-rank group
    ref.ep_group = solo[rank]
    assert ep.experts.w_up.shape[0] == cfg.moe.num_experts // world_size
    torch.manual_seed(5 + rank)
    x = torch.randn(2, 16, cfg.hidden_size)
    xr = x.clone().requires_grad_()
    xe = x.clone().requires_grad_()
    yr = ref(xr)
    ye = ep(xe)
    torch.testing.assert_close(ye, yr, rtol=1e-4, atol=1e-5)
    yr.square().sum().backward()
    ye.square().sum().backward()
    torch.testing.assert_close(xe.grad, xr.grad, rtol=1e-4, atol=1e-5)
    dist.barrier()
    dist.destroy_process_group()


@rerun_if_address_is_in_use()
def test_expert_parallel_matches_single_rank_gloo():
    spawn(_ep_worker, 2)


def test_padded_expert_layout_matches_packed_layout_cpu(monkeypatch):
    """The 128-row padded expert layout of the sorted path (used on the GPU so that the grouped weight-gradient GEMM
    needs no re-layout) is an index transformation only: forced on for CPU tensors it must reproduce the packed path
    bit for bit - output and all gradients - including experts that receive no token."""
    from colossalai_b200.moe.grouped_gemm import _pad_groups, grouped_linear

    torch.manual_seed(0)
    T, H, E, K = 50, 16, 6, 2
    w = torch.randn(E, H, H, requires_grad=True)
    x = torch.randn(T, H, requires_grad=True)
    lg = torch.randn(T, E)
    lg[:, 4] = -1e9                                         # expert 4 never selected
    lg.requires_grad_()

    def run(aligned):
        monkeypatch.setattr(dc, "_aligned_rows", lambda t: aligned)
        for t in (w, x, lg):
            t.grad = None
        tw, ti = lg.softmax(-1).topk(K, -1)
        y = dc.moe_forward(x, tw, ti, lambda r, c: grouped_linear(r, w, c), E, None)
        y.square().sum().backward()
        return y.detach().clone(), x.grad.clone(), lg.grad.clone(), w.grad.clone()

    for a, b in zip(run(False), run(True)):
        torch.testing.assert_close(a, b, rtol=0, atol=0)
    # layout helper: every group starts on a multiple of 128, rows keep their order inside a group, rows of an
    # over-allocated buffer behind the last group go to the dump row
    counts = torch.tensor([3, 0, 130, 1])
    dest, pends, bound = _pad_groups(counts, 140, "cpu")
    assert pends.tolist() == [128, 128, 384, 512] and bound % 128 == 0 and bound >= 512
    assert dest[:3].tolist() == [0, 1, 2] and dest[3:133].tolist() == list(range(128, 258)) and dest[133].item() == 384
    assert (dest[134:] == bound).all()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_legacy_dispatch_combine_gpu(dtype):
    _legacy_roundtrip("cuda", dtype)


@pytest.mark.gpu
def test_cumsum_and_router_gpu():
    torch.manual_seed(0)
    mask = (torch.rand(3000, 8, device="cuda") < 0.3).to(torch.int32)
    torch.testing.assert_close(moe_cumsum(mask), (torch.cumsum(mask, 0) - 1).to(torch.int32))
    for E, k, dtype in [(8, 2, torch.bfloat16), (64, 6, torch.float32), (160, 8, torch.float32)]:
        lg = torch.randn(777, E, device="cuda").to(dtype)
        w, idx, probs = mops.router_topk(lg, k, return_probs=True)
        p = lg.float().softmax(-1)
        rw, ridx = p.topk(k, -1)
        torch.testing.assert_close(probs, p, rtol=1e-4, atol=1e-6)
        # equal probabilities (bf16 ties) may be picked in a different order: compare the selected VALUES
        torch.testing.assert_close(p.gather(1, idx).sort(-1).values, rw.sort(-1).values, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(w.sum(-1), torch.ones(777, device="cuda"), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(w, p.gather(1, idx) / p.gather(1, idx).sum(-1, keepdim=True), rtol=1e-4, atol=1e-6)


def _fused_vs_nccl(group, T, H, E, K, seed):
    from colossalai_b200.moe.grouped_gemm import grouped_linear

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_local = E // world
    torch.manual_seed(100)
    w_all = (torch.randn(E, H, H, device="cuda") * 0.05).bfloat16()
    w_loc = w_all[rank * n_local:(rank + 1) * n_local].clone().requires_grad_()
    w_loc2 = w_loc.detach().clone().requires_grad_()

    def experts_for(w):
        return lambda rows, counts: grouped_linear(rows, w, counts)

    torch.manual_seed(seed + rank)
    x = torch.randn(T, H, device="cuda").bfloat16()
    lg = torch.randn(T, E, device="cuda")
    outs = {}
    for name, w in (("nccl", w_loc), ("fused", w_loc2)):
        dc.set_moe_backend(name)
        xi = x.clone().requires_grad_()
        li = lg.clone().requires_grad_()
        p = li.softmax(-1)
        tw, ti = p.topk(K, -1)
        y = dc.moe_forward(xi, tw, ti, experts_for(w), E, group)
        (y.float() * torch.linspace(-1, 1, H, device="cuda")).sum().backward()
        outs[name] = (y.detach().float(), xi.grad.float(), li.grad.float(), w.grad.float())
    dc.set_moe_backend("auto")
    for a, b, what in zip(outs["fused"], outs["nccl"], ("y", "dx", "dlogits", "dw")):
        # rows reach the experts in a different order -> bf16 outputs may differ by one ulp: compare in norm
        err = (a - b).norm() / b.norm().clamp_min(1e-6)
        assert err < 1e-2, f"{what}: relative error {err:.3e}"
        assert (a - b).abs().max() <= 0.02 * b.abs().max() + 1e-3, f"{what}: max abs diff {(a - b).abs().max():.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("T,H,E,K", [(300, 256, 8, 2), (1000, 128, 16, 4), (64, 128, 4, 1)])
def test_moe_forward_gpu_matches_dense_reference(T, H, E, K):
    """Sorted (no expert parallelism) path on the GPU - expert rows in the 128-padded layout, native grouped GEMMs -
    against a dense fp32 evaluation of every (token, selected expert) pair: output, dx, dlogits, dW."""
    from colossalai_b200.moe.grouped_gemm import grouped_linear

    torch.manual_seed(5)
    w = (torch.randn(E, H, H, device="cuda") * 0.05).bfloat16().requires_grad_()
    x = torch.randn(T, H, device="cuda").bfloat16().requires_grad_()
    lg = torch.randn(T, E, device="cuda", requires_grad=True)
    dc.set_moe_backend("nccl")
    try:
        tw, ti = lg.softmax(-1).topk(K, -1)
        y = dc.moe_forward(x, tw, ti, lambda rows, counts: grouped_linear(rows, w, counts), E, None)
        proj = torch.linspace(-1, 1, H, device="cuda")
        (y.float() * proj).sum().backward()
    finally:
        dc.set_moe_backend("auto")
    got = (y.detach().float(), x.grad.float(), lg.grad.float(), w.grad.float())
    xr = x.detach().float().requires_grad_()
    wr = w.detach().float().requires_grad_()
    lr = lg.detach().clone().requires_grad_()
    twr, tir = lr.softmax(-1).topk(K, -1)
    all_out = torch.einsum("th,eoh->teo", xr, wr)                      # every token through every expert
    yr = (all_out.gather(1, tir[:, :, None].expand(-1, -1, H)) * twr[:, :, None]).sum(1)
    (yr * proj).sum().backward()
    for a, b, what in zip(got, (yr.detach(), xr.grad, lr.grad, wr.grad), ("y", "dx", "dlogits", "dw")):
        err = (a - b).norm() / b.norm().clamp_min(1e-6)
        assert err < 2e-2, f"{what}: relative error {err:.3e}"


@pytest.mark.gpu
def test_fused_ep_single_gpu_matches_sorted_path():
    _fused_vs_nccl(None, T=300, H=256, E=8, K=2, seed=1)
    _fused_vs_nccl(None, T=64, H=128, E=4, K=1, seed=2)     # buffers are reused across calls (epochs)


def _fused_ep_worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="nccl", verbose=False)
    for i in range(3):
        _fused_vs_nccl(dist.group.WORLD, T=512 + 64 * rank, H=512, E=8, K=2, seed=10 + i)
    dist.barrier()
    if rank == 0:
        print("FUSED_EP_OK", flush=True)
    dist.destroy_process_group()


@pytest.mark.gpu
@rerun_if_address_is_in_use()
def test_fused_ep_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    spawn(_fused_ep_worker, 2)


if __name__ == "__main__":
    import os

    spawn(_fused_ep_worker, int(os.environ.get("NGPU", "2")))
