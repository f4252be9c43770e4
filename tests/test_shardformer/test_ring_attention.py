"""Zigzag ring attention == full causal attention (values and gradients).  CPU tier on gloo; GPU tier exercises the
library flash kernel with LSE + the P2P KV gather / fused reduce-scatter path (reference:
tests/test_shardformer/test_layer/test_ring_attn.py)."""
import os

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.ops.attention import attention_ref
from colossalai_b200.shardformer.layer._operation import split_batch_zigzag
from colossalai_b200.shardformer.layer.attn import RingAttention
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _check(device, dtype, B, S, Hq, Hkv, D, tol):
    sp, r = dist.get_world_size(), dist.get_rank()
    torch.manual_seed(0)
    q = torch.randn(B, S, Hq, D, device=device, dtype=dtype)
    k = torch.randn(B, S, Hkv, D, device=device, dtype=dtype)
    v = torch.randn(B, S, Hkv, D, device=device, dtype=dtype)
    do = torch.randn(B, S, Hq, D, device=device, dtype=dtype)
    # oracle on the full sequence (fp32)
    qf, kf, vf = (t.float().reshape(B * S, t.shape[2], D).requires_grad_() for t in (q, k, v))
    ref = attention_ref(qf, kf, vf, batch=B, causal=True)
    ref.backward(do.float().reshape(B * S, Hq, D))
    # local zigzag shards
    ql, kl, vl, dol = (split_batch_zigzag(t, dist.group.WORLD, seq_dim=1).contiguous() for t in (q, k, v, do))
    Sl = S // sp
    ql, kl, vl = (t.reshape(B * Sl, t.shape[2], D).requires_grad_() for t in (ql, kl, vl))
    out = RingAttention.attention(ql, kl, vl, dist.group.WORLD, batch=B)
    out.backward(dol.reshape(B * Sl, Hq, D))

    def shard(full, H):
        return split_batch_zigzag(full.view(B, S, H, D), dist.group.WORLD, seq_dim=1).reshape(B * Sl, H, D)

    torch.testing.assert_close(out.float(), shard(ref.detach(), Hq), **tol)
    torch.testing.assert_close(ql.grad.float(), shard(qf.grad, Hq), **tol)
    torch.testing.assert_close(kl.grad.float(), shard(kf.grad, Hkv), **tol)
    torch.testing.assert_close(vl.grad.float(), shard(vf.grad, Hkv), **tol)


def _cpu_worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    _check("cpu", torch.float32, B=2, S=32, Hq=4, Hkv=2, D=16, tol=dict(rtol=1e-4, atol=1e-5))
    dist.barrier()
    dist.destroy_process_group()


@rerun_if_address_is_in_use()
def test_ring_attention_cpu():
    spawn(_cpu_worker, 2)


def _gpu_worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="nccl", verbose=False)
    from colossalai_b200.parallel import fused

    tol = dict(rtol=3e-2, atol=3e-2)
    _check("cuda", torch.bfloat16, B=2, S=1024, Hq=8, Hkv=2, D=128, tol=tol)      # P2P gather + fused RS path
    assert fused.stats.get("reduce_scatter", 0) > 0 and fused.stats["all_gather"] > 0, fused.stats
    os.environ["CB200_RING_ATTN_P2P"] = "0"
    _check("cuda", torch.bfloat16, B=1, S=2048, Hq=8, Hkv=8, D=64, tol=tol)       # NCCL ring path
    # stand-alone fused reduce-scatter, both element types
    for dt in (torch.float32, torch.bfloat16):
        torch.manual_seed(5 + rank)
        x = torch.randn(world_size * 300, 64, device="cuda").to(dt)
        ref = x.float().clone()
        dist.all_reduce(ref)
        got = fused.reduce_scatter(x, dist.group.WORLD)
        torch.testing.assert_close(got.float(), ref[rank * 300:(rank + 1) * 300], rtol=2e-2, atol=2e-2)
    dist.barrier()
    if rank == 0:
        print("RING_ATTN_GPU_OK", flush=True)
    dist.destroy_process_group()


@pytest.mark.gpu
@rerun_if_address_is_in_use()
def test_ring_attention_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    spawn(_gpu_worker, 2)


if __name__ == "__main__":
    spawn(_gpu_worker, int(os.environ.get("NGPU", "2")))
