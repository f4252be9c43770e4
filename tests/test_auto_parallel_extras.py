"""Rotor checkpoint solver, offload planner/runtime, AutoChunk and the fx split / checkpoint passes (reference:
tests/test_auto_parallel/test_ckpt_solvers, test_offload, tests/test_autochunk, tests/test_fx/test_pipeline)."""
import copy

import pytest
import torch
import torch.nn as nn

from colossalai_b200.auto_parallel.autochunk import ChunkedModule, autochunk, is_chunkable
from colossalai_b200.auto_parallel.checkpoint import Chain, CheckpointSolverRotor, apply_rotor_checkpointing
from colossalai_b200.auto_parallel.offload import AsynGreedySolver, SynGreedySolver, memory_optimize
from colossalai_b200.auto_parallel.offload.solver import LayerCost
from colossalai_b200.fx import symbolic_trace
from colossalai_b200.fx.passes import (activation_checkpoint_pass, balanced_split_pass, split_with_split_nodes_pass,
                                       uniform_split_pass)


def test_rotor_time_memory_tradeoff():
    n = 8
    ch = Chain([1.0] * n, [2.0] * n, [1.0] * (n + 1), [4.0] * n)
    times = []
    for budget in (80, 20, 12, 8):
        s = CheckpointSolverRotor(ch, budget, 128)
        seq = s.solve()
        assert abs(seq.makespan(ch) - s.optimal_time()) < 1e-9
        assert sorted(i for op, i in seq.ops if op == "B") == list(range(n))       # every stage back-propagated once
        times.append(s.optimal_time())
    assert times[0] == 3.0 * n                      # enough memory: no recomputation
    assert times == sorted(times) and times[-1] > times[0]
    with pytest.raises(RuntimeError):
        CheckpointSolverRotor(ch, 3, 64).solve()


def test_rotor_applied_to_a_model_keeps_gradients():
    torch.manual_seed(0)
    stages = [nn.Sequential(nn.Linear(32, 128), nn.GELU(), nn.Linear(128, 32)) for _ in range(6)]
    x = torch.randn(16, 32)
    ref = nn.Sequential(*copy.deepcopy(stages))
    full_bytes = 6 * (16 * 128 * 4 * 3) + 7 * (16 * 32 * 4)      # ~ saved activations of all stages + stage outputs
    model, seq, chain = apply_rotor_checkpointing(stages, x, memory_budget=0.45 * full_bytes)
    assert seq.checkpoint_segments(), "a tight budget must trigger recomputation"
    model.train(), ref.train()
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    model(xa).sum().backward()
    ref(xb).sum().backward()
    torch.testing.assert_close(xa.grad, xb.grad)


def test_offload_solvers_and_runtime():
    layers = [LayerCost(f"l{i}", 100.0, 1e-9 * (i + 1)) for i in range(6)]
    syn = SynGreedySolver(layers, budget_bytes=350.0, bandwidth=1e9).solve()
    asy = AsynGreedySolver(layers, budget_bytes=350.0, bandwidth=1e9, prefetch_distance=1).solve()
    assert syn.peak_param_bytes <= 350 and asy.peak_param_bytes <= 350 and len(syn.offloaded) >= 3
    assert asy.est_stall_s <= syn.est_stall_s * len(asy.offloaded) / max(1, len(syn.offloaded)) + 1e-12
    assert all(asy.prefetch_at[n] != n or n == "l0" for n in asy.offloaded)
    with pytest.raises(RuntimeError):
        SynGreedySolver(layers, budget_bytes=50.0).solve()
    # runtime: forward/backward still correct with half of the blocks living on the host
    torch.manual_seed(0)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.blocks = nn.ModuleList([nn.Linear(32, 32) for _ in range(6)])

        def forward(self, x):
            for b in self.blocks:
                x = torch.tanh(b(x))
            return x

    net, ref = Net(), None
    ref = copy.deepcopy(net)
    total = sum(p.numel() * 4 for p in net.parameters())
    net, plan = memory_optimize(net, budget_bytes=0.6 * total, solver_name="asyn")
    assert plan.offloaded and all(net.blocks[int(n.split(".")[1])].weight.numel() == 0 for n in plan.offloaded)
    x = torch.randn(4, 32)
    with torch.no_grad():
        torch.testing.assert_close(net(x), ref(x))
    assert all(net.blocks[int(n.split(".")[1])].weight.numel() == 0 for n in plan.offloaded)   # released again
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    net(xa).sum().backward()
    ref(xb).sum().backward()
    torch.testing.assert_close(xa.grad, xb.grad)


def test_autochunk_wraps_only_what_exceeds_the_budget():
    from colossalai_b200.models import build_model
    from colossalai_b200.models.transformer import MLP

    torch.manual_seed(0)
    model = build_model("llama-tiny").float()
    ids = torch.randint(0, 512, (2, 64))
    ref = model(input_ids=ids, labels=ids)["loss"]
    assert is_chunkable(model.model.layers[0].mlp, torch.randn(128, 64), 0)
    assert not is_chunkable(nn.Sequential(nn.Linear(8, 8), nn.Softmax(dim=0)), torch.randn(6, 8), 0)
    model, regions = autochunk(model, dict(input_ids=ids), max_memory_bytes=64 * 1024, dim=0)
    assert len(regions) == model.cfg.num_hidden_layers and all(n >= 2 for _, _, n in regions)
    assert isinstance(model.model.layers[0].mlp, ChunkedModule)
    out = model(input_ids=ids, labels=ids)["loss"]
    torch.testing.assert_close(out, ref)
    out.backward()
    m2, r2 = autochunk(build_model("llama-tiny").float(), dict(input_ids=ids), max_memory_bytes=1 << 30)
    assert not r2 and isinstance(m2.model.layers[0].mlp, MLP)


def test_fx_split_and_checkpoint_passes():
    torch.manual_seed(0)
    net = nn.Sequential(*[nn.Sequential(nn.Linear(16, 16), nn.ReLU()) for _ in range(4)], nn.Linear(16, 4))
    gm = symbolic_trace(net)
    x = torch.randn(3, 16)
    for pass_fn in (balanced_split_pass, uniform_split_pass):
        part = pass_fn(gm, 2)
        split, stages = split_with_split_nodes_pass(gm, part)
        assert len(stages) == 2
        torch.testing.assert_close(split(x), net(x))
        torch.testing.assert_close(stages[1](stages[0](x)), net(x))
    names = [n.name for n in gm.graph.nodes if n.op == "call_module"]
    ck = activation_checkpoint_pass(gm, [names[0:4], names[4:8]])
    from colossalai_b200.fx.passes import CheckpointRegion

    assert sum(isinstance(m, CheckpointRegion) for m in ck.children()) == 2
    ck.train()
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    ck(xa).sum().backward()
    net(xb).sum().backward()
    torch.testing.assert_close(xa.grad, xb.grad)


def test_analyzer_meta_tensor_flop_count_and_symbolic_profile():
    """reference tests/test_analyzer/{test_subclasses,test_fx}: device-free tensors, fwd/bwd FLOPs, per-node MetaInfo."""
    import torch.nn as nn

    from colossalai_b200._analyzer import MeshConfig, MetaTensor, MetaTensorMode, flop_count
    from colossalai_b200._analyzer.fx import symbolic_profile, symbolic_trace

    class MLP(nn.Module):
        def __init__(self, h=64, f=256):
            super().__init__()
            self.up, self.act, self.down = nn.Linear(h, f), nn.GELU(), nn.Linear(f, h)

        def forward(self, x):
            return self.down(self.act(self.up(x)))

    m = MLP()
    fwd, bwd = flop_count(m, torch.randn(8, 64))
    gemm = 2 * 8 * 64 * 256 * 2
    assert gemm <= fwd < 1.05 * gemm                      # GEMMs + a little pointwise work
    assert 1.9 * gemm < bwd < 2.1 * gemm                  # dgrad + wgrad of both layers
    with MetaTensorMode():
        big = nn.Linear(8192, 4 * 8192)                   # 1 GiB of weights, nothing allocated
        x = torch.randn(4, 8192, device="cuda:0")
    assert isinstance(x, MetaTensor) and x.device == torch.device("cuda:0") and x._tensor.device.type == "meta"
    y = big(x)
    assert isinstance(y, MetaTensor) and tuple(y.shape) == (4, 4 * 8192)
    t = MetaTensor(torch.randn(3, 4), device="cuda:1")
    assert (t @ t.t()).device == torch.device("cuda:1") and t.to("cpu").device.type == "cpu"
    gm = symbolic_profile(symbolic_trace(m), torch.randn(8, 64))
    info = {n.name: n.meta["info"] for n in gm.graph.nodes}
    assert info["up"].fwd_flop >= 2 * 8 * 64 * 256 and info["up"].bwd_flop >= 2 * info["up"].fwd_flop - 8 * 256 * 2
    assert info["up"].param_bytes == (64 * 256 + 256) * 4 and info["act"].saved_bytes == 8 * 256 * 4
    assert info["down"].outputs == ((8, 64), torch.float32)
    cfg = MeshConfig()
    assert cfg.TFLOPS > 1000 and cfg.HBM_BYTES == 180e9
