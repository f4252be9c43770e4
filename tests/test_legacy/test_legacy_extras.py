"""Legacy API extras: ParallelMode-keyed communication, 1-D layers, ring sequence parallelism (RingQK / RingAV),
capacity-based MoE with expert parallelism and the load balancer, first-generation ZeRO adapters (reference:
tests/test_legacy/{test_comm,test_layers/test_1d,test_layers/test_sequence,test_moe,test_zero})."""
import copy
import math

import pytest
import torch
import torch.distributed as dist
import torch.nn.functional as F

import colossalai_b200
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _setup(rank, world_size, port):
    from colossalai_b200.legacy.context import ParallelMode, global_context as gpc

    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    g = dist.new_group(list(range(world_size)))
    for mode in (ParallelMode.PARALLEL_1D, ParallelMode.SEQUENCE, ParallelMode.PIPELINE, ParallelMode.TENSOR):
        gpc.set_group(mode, g)
    return gpc, ParallelMode, g


def _comm_and_layers(rank, world_size, port):
    gpc, PM, g = _setup(rank, world_size, port)
    from colossalai_b200.legacy import communication as C
    from colossalai_b200.legacy.nn.layer import Linear1D, TransformerSelfAttentionRing

    # collectives
    x = torch.full((2, 3), float(rank + 1))
    assert C.all_gather(x, 0, PM.PARALLEL_1D).shape == (4, 3)
    assert torch.equal(C.all_reduce(x.clone(), PM.PARALLEL_1D), torch.full((2, 3), 3.0))
    rs = C.reduce_scatter(torch.arange(8.0).view(4, 2), 0, PM.PARALLEL_1D)
    assert torch.equal(rs, 2 * torch.arange(8.0).view(4, 2)[rank * 2:(rank + 1) * 2])
    got = C.ring_forward(torch.tensor([float(rank)]), PM.SEQUENCE)
    assert got.item() == float((rank - 1) % world_size)
    # pipeline p2p
    if rank == 0:
        C.send_forward(torch.arange(6.0).view(2, 3))
        gback = C.recv_backward((2, 3))
        assert torch.equal(gback, torch.ones(2, 3) * 5)
    else:
        act = C.recv_forward((2, 3))
        assert torch.equal(act, torch.arange(6.0).view(2, 3))
        C.send_backward(torch.ones(2, 3) * 5)
    # Linear1D: widen (col) then narrow (row) == dense MLP
    torch.manual_seed(3)
    l1, l2 = Linear1D(8, 16), Linear1D(16, 8)
    from colossalai_b200.shardformer.layer.linear import Linear1D_Col, Linear1D_Row

    assert isinstance(l1.layer, Linear1D_Col) and isinstance(l2.layer, Linear1D_Row)
    from colossalai_b200.parallel import comm

    w1 = comm.all_gather(l1.weight.data, 0, g)
    b1 = comm.all_gather(l1.bias.data, 0, g)
    w2 = comm.all_gather(l2.weight.data, 1, g)
    xin = torch.randn(5, 8, generator=torch.Generator().manual_seed(9))
    ref = F.linear(torch.relu(F.linear(xin, w1, b1)), w2, l2.bias.data)
    torch.testing.assert_close(l2(torch.relu(l1(xin))), ref, atol=1e-5, rtol=1e-5)

    # ring sequence parallel attention == dense attention on the full sequence
    torch.manual_seed(11)
    attn = TransformerSelfAttentionRing(16, 4, causal=True)
    for p in attn.parameters():
        dist.broadcast(p.data, 0)
    S, B = 8, 2
    full = torch.randn(S, B, 16, generator=torch.Generator().manual_seed(5))
    local = full[rank * (S // world_size):(rank + 1) * (S // world_size)].clone().requires_grad_()
    out = attn(local)
    out.sum().backward()
    # dense oracle
    fx = full.clone().requires_grad_()
    qkv = F.linear(fx, attn.query_key_value.weight.detach(), attn.query_key_value.bias.detach())
    qkv = qkv.view(S, B * 4, 12).transpose(0, 1)
    q, k, v = qkv.chunk(3, -1)
    sc = q @ k.transpose(1, 2) / math.sqrt(4)
    sc = sc.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), float("-inf"))
    o = (sc.softmax(-1) @ v).transpose(0, 1).reshape(S, B, 16)
    o = F.linear(o, attn.dense.weight.detach(), attn.dense.bias.detach())
    sl = slice(rank * (S // world_size), (rank + 1) * (S // world_size))
    torch.testing.assert_close(out, o[sl], atol=1e-5, rtol=1e-4)
    o.sum().backward()
    torch.testing.assert_close(local.grad, fx.grad[sl], atol=1e-5, rtol=1e-4)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_legacy_comm_1d_and_ring_sequence_parallel():
    spawn(_comm_and_layers, 2)


def test_legacy_moe_kernel_matches_einsum_and_balancer():
    from colossalai_b200.legacy.moe import MOE_MANAGER, LoadBalancer, SparseMLP, Top1Router

    MOE_MANAGER.setup(parallel=None)
    MOE_MANAGER.reset_loss()
    torch.manual_seed(0)
    m = SparseMLP(4, 16, 32, router_top_k=2, router_capacity_factor_train=4.0)
    mk = SparseMLP(4, 16, 32, router_top_k=2, router_capacity_factor_train=4.0, enable_kernel=True)
    mk.load_state_dict(m.state_dict())
    x = torch.randn(2, 10, 16)
    torch.testing.assert_close(m(x), mk(x), atol=1e-6, rtol=1e-5)
    aux, z = MOE_MANAGER.get_loss()
    assert len(aux) == 2 and float(aux[0]) > 0
    # capacity: with a tiny capacity tokens are dropped (output rows are exactly zero)
    r = Top1Router(capacity_factor_train=0.25, min_capacity=1)
    used, combine, sec = r(torch.randn(32, 4))
    assert sec.shape[-1] == r.get_capacity(32, 4) and int(sec.sum()) <= 4 * sec.shape[-1]
    # balancer moves a hot expert away from the overloaded rank
    lb = LoadBalancer(m.experts, m.gate_weight, 2, 4)
    lb.update_load(torch.tensor([100.0, 90.0, 5.0, 5.0]))
    before = lb._rank_loads(lb.placement, lb.local_load)
    placement, swaps = lb._search_balance(lb.placement, lb.local_load)
    after = lb._rank_loads(placement, lb.local_load)
    assert swaps and max(after) - min(after) < max(before) - min(before)
    w0 = m.experts.wi.data.clone()
    lb.placement = [[0, 1, 2, 3]]                  # single-rank placement: swaps happen inside the local stack
    lb._swap_moe_param([((0, 0), (0, 2))])
    assert torch.equal(m.experts.wi.data[0], w0[2]) and torch.equal(m.experts.wi.data[2], w0[0])


def _moe_ep(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    from colossalai_b200.legacy.moe import MOE_MANAGER, SparseMLP

    MOE_MANAGER.setup(parallel=None)
    torch.manual_seed(0)
    dense = SparseMLP(4, 16, 32, router_top_k=1, router_capacity_factor_train=8.0)
    MOE_MANAGER.setup(parallel="EP", ep_size=2)
    ep = SparseMLP(4, 16, 32, router_top_k=1, router_capacity_factor_train=8.0)
    ep.gate_weight.data.copy_(dense.gate_weight.data)
    ep.experts.wi.data.copy_(dense.experts.wi.data[rank * 2:(rank + 1) * 2])
    ep.experts.wo.data.copy_(dense.experts.wo.data[rank * 2:(rank + 1) * 2])
    x = torch.randn(3, 6, 16, generator=torch.Generator().manual_seed(7 + rank), requires_grad=True)
    xr = x.detach().clone().requires_grad_()
    y, yr = ep(x), dense(xr)
    torch.testing.assert_close(y, yr, atol=1e-5, rtol=1e-5)
    y.sum().backward()
    yr.sum().backward()
    torch.testing.assert_close(x.grad, xr.grad, atol=1e-5, rtol=1e-5)
    MOE_MANAGER.setup(parallel=None)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_legacy_moe_expert_parallel():
    spawn(_moe_ep, 2)


def _zero_v2(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    from colossalai_b200.legacy.zero import (BucketTensorShardStrategy, ShardedModelV2, ShardedOptimizerV2,
                                             ZeroInitContext)

    numel = torch.zeros(1, dtype=torch.long)
    with ZeroInitContext(target_device=torch.device("cpu"), shard_strategy=BucketTensorShardStrategy(),
                         shard_param=True, bf16=True, model_numel_tensor=numel):
        net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 4))
    assert int(numel) == 16 * 32 + 32 + 32 * 4 + 4 and net[0].weight.dtype == torch.bfloat16
    model = ShardedModelV2(net, BucketTensorShardStrategy(), bf16=True, tensor_placement_policy="cuda")
    optim = ShardedOptimizerV2(model, torch.optim.Adam(model.parameters(), lr=1e-2), initial_scale=1)
    x = torch.randn(8, 16, generator=torch.Generator().manual_seed(rank))
    losses = []
    for _ in range(3):
        loss = model(x).float().pow(2).mean()
        optim.backward(loss)
        optim.step()
        optim.zero_grad()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_legacy_zero_v2_adapters():
    spawn(_zero_v2, 2)


def test_registry_builder_checkpoint_sampler_profilers():
    from colossalai_b200.legacy.builder import build_from_registry
    from colossalai_b200.legacy.registry import LAYERS, OPTIMIZERS, Registry
    from colossalai_b200.legacy.utils import DataParallelSampler, checkpoint, clip_grad_norm_fp32
    from colossalai_b200.legacy.utils.profiler import MemProfiler, PcieProfiler, ProfilerContext
    from colossalai_b200.utils.profiler import CommProfiler

    reg = Registry("toy")

    @reg.register_module
    class Foo:
        def __init__(self, a=1):
            self.a = a

    assert build_from_registry(dict(type="Foo", a=3), reg).a == 3 and LAYERS.has("Linear")
    lin = build_from_registry(dict(type="Linear", in_features=4, out_features=2), LAYERS)
    opt = OPTIMIZERS.get_module("SGD")(lin.parameters(), lr=0.1)
    assert isinstance(opt, torch.optim.SGD)
    # activation checkpoint with offload and dropout RNG replay
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Dropout(0.5), torch.nn.Linear(8, 8))
    x = torch.randn(4, 8, requires_grad=True)
    torch.manual_seed(5)
    y = checkpoint(net, True, x)
    y.sum().backward()
    g1 = x.grad.clone()
    x.grad = None
    torch.manual_seed(5)
    net(x).sum().backward()
    torch.testing.assert_close(g1, x.grad)
    # sampler: 2 replicas would partition, single replica sees everything exactly once
    s = DataParallelSampler(list(range(10)), shuffle=True, seed=3)
    assert sorted(iter(s)) == list(range(10)) and len(s) == 10
    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.full((4,), 3.0)
    n = clip_grad_norm_fp32([p], max_norm=1.0)
    assert abs(float(n) - 6.0) < 1e-5 and abs(float(p.grad.norm()) - 1.0) < 1e-3
    with ProfilerContext([CommProfiler(), PcieProfiler(), MemProfiler()]) as prof:
        prof.profilers[0].step("a") if hasattr(prof.profilers[0], "step") else None
        torch.randn(8, 8) @ torch.randn(8, 8)
    assert "Pcie profiling result" in "".join(p.result_str() for p in prof.profilers)


def _legacy_init(rank, world_size, port):
    from colossalai_b200.legacy.amp import AMP_TYPE
    from colossalai_b200.legacy.context import ParallelMode, global_context as gpc
    from colossalai_b200.legacy.initialize import initialize, launch

    launch(dict(parallel=dict(pipeline=1, tensor=dict(size=1, mode="1d")), clip_grad_norm=1.0,
                fp16=dict(mode=AMP_TYPE.NAIVE, dtype=torch.bfloat16)),
           rank=rank, world_size=world_size, host="127.0.0.1", port=port, backend="gloo", verbose=False)
    assert gpc.get_world_size(ParallelMode.DATA) == 2 and gpc.get_world_size(ParallelMode.TENSOR) == 1
    torch.manual_seed(rank)            # replicas start different; initialize() must sync them
    model = torch.nn.Linear(8, 2)
    engine, *_ = initialize(model, torch.optim.SGD(model.parameters(), lr=0.1), torch.nn.MSELoss())
    w = engine.model.weight.detach().float().clone()
    dist.all_reduce(w)
    torch.testing.assert_close(w / 2, engine.model.weight.detach().float())
    engine.train()
    x = torch.randn(4, 8, generator=torch.Generator().manual_seed(rank)).bfloat16()
    loss = engine.criterion(engine(x).float(), torch.zeros(4, 2))
    engine.zero_grad()
    engine.optimizer.backward(loss) if hasattr(engine.optimizer, "backward") else engine.backward(loss)
    engine.step()
    w2 = engine.model.weight.detach().float().clone()
    dist.all_reduce(w2)
    torch.testing.assert_close(w2 / 2, engine.model.weight.detach().float())     # replicas stay in lock-step
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_legacy_initialize_engine():
    spawn(_legacy_init, 2)
