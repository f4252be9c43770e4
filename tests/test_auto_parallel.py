"""fx tracing / meta profiling and the chain sharding solver (reference: tests/test_fx, tests/test_auto_parallel)."""
import torch
import torch.nn as nn

from colossalai_b200.auto_parallel import initialize_model
from colossalai_b200.device import DeviceMesh
from colossalai_b200.fx import MetaInfoProp, profile_flops_and_memory, symbolic_trace


class MLP(nn.Module):
    def __init__(self, h=64, f=256):
        super().__init__()
        self.up = nn.Linear(h, f)
        self.act = nn.GELU()
        self.down = nn.Linear(f, h)

    def forward(self, x):
        return self.down(self.act(self.up(x)))


def test_trace_and_meta_prop():
    m = MLP()
    gm = symbolic_trace(m, meta_args={"x": torch.empty(8, 64, device="meta")})
    x = torch.randn(8, 64)
    torch.testing.assert_close(gm(x), m(x))
    prop = MetaInfoProp(gm)
    prop.propagate(torch.empty(8, 64))
    by_name = {n.name: n for n in gm.graph.nodes}
    assert by_name["up"].meta["tensor_meta"][0] == (8, 256)
    assert by_name["up"].meta["fwd_flop"] == 2 * 8 * 64 * 256
    assert "GFLOP" in prop.summary()
    flops, acts = profile_flops_and_memory(m, torch.empty(8, 64))
    assert flops == 2 * 8 * 64 * 256 * 2 and acts > 0


def test_chain_solver_prefers_col_then_row_and_respects_memory():
    mesh = DeviceMesh(torch.arange(8), (1, 8))
    big = MLP(h=4096, f=14336)
    meta = {"x": torch.empty(8, 4096, 1, device="meta")}          # 32768 tokens
    plan = initialize_model(big, meta, mesh)
    assert plan.as_dict() == {"up": "col", "down": "row"}           # Megatron pairing: one all-reduce, no gather
    rep_bytes = 2 * 4096 * 14336 * 2
    assert plan.param_bytes_per_device < rep_bytes / 4
    tiny = MLP(h=8, f=16)
    plan_t = initialize_model(tiny, {"x": torch.empty(2, 4, 1, device="meta")}, mesh)
    assert set(plan_t.as_dict().values()) == {"replicate"}          # latency dominates: do not shard
    forced = initialize_model(tiny, {"x": torch.empty(2, 4, 1, device="meta")}, mesh, memory_budget=300)
    assert forced.param_bytes_per_device <= 300
