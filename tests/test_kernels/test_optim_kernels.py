"""Fused / CPU Adam vs torch reference over the dtype matrix (reference: tests/test_optimizer/test_adam_kernel.py)."""
import pytest
import torch

from colossalai_b200.nn.optimizer import CPUAdam, FusedAdam, HybridAdam
from colossalai_b200.nn.optimizer.cpu_adam import cpu_adam_step
from colossalai_b200.nn.optimizer.fused_adam import adam_reference_step


@pytest.mark.parametrize("adamw", [True, False])
@pytest.mark.parametrize("p_dtype,g_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                             (torch.float32, torch.bfloat16), (torch.float16, torch.float16)])
def test_cpu_adam_kernel(adamw, p_dtype, g_dtype):
    torch.manual_seed(0)
    n = 100003
    p = torch.randn(n).to(p_dtype)
    g = torch.randn(n).to(g_dtype)
    m, v = torch.zeros(n), torch.zeros(n)
    pr, mr, vr = p.clone().float(), m.clone(), v.clone()
    for step in range(1, 4):
        cpu_adam_step(p, g, m, v, 1e-2, 0.9, 0.999, 1e-8, 0.05, step, True, adamw)
        adam_reference_step([pr], [g.float()], [mr], [vr], 1e-2, 0.9, 0.999, 1e-8, 0.05, step, adamw, True)
        if p_dtype != torch.float32:   # the low-precision param is re-quantised every step: mirror it
            pr = pr.to(p_dtype).float()
    tol = dict(atol=1e-5, rtol=1e-4) if p_dtype == torch.float32 else dict(atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(p.float(), pr, **tol)
    torch.testing.assert_close(m, mr, atol=1e-5, rtol=1e-4)


def test_cpu_adam_optimizer_matches_torch():
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(1000, 33))
    w2 = torch.nn.Parameter(w.detach().clone())
    o1, o2 = CPUAdam([w], lr=1e-2, weight_decay=0.1), torch.optim.AdamW([w2], lr=1e-2, weight_decay=0.1)
    for _ in range(3):
        g = torch.randn_like(w)
        w.grad, w2.grad = g.clone(), g.clone()
        o1.step()
        o2.step()
    torch.testing.assert_close(w, w2, atol=1e-5, rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("adamw", [True, False])
@pytest.mark.parametrize("p_dtype,g_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                             (torch.float16, torch.float32)])
def test_fused_adam_kernel(adamw, p_dtype, g_dtype):
    torch.manual_seed(0)
    shapes = [(1024, 1024), (333,), (7, 4099), (2048 * 16 + 5,)]
    ps = [torch.nn.Parameter(torch.randn(*s, device="cuda").to(p_dtype)) for s in shapes]
    refs = [p.detach().clone().float() for p in ps]
    ms = [torch.zeros_like(r) for r in refs]
    vs = [torch.zeros_like(r) for r in refs]
    opt = FusedAdam(ps, lr=1e-2, weight_decay=0.05, adamw_mode=adamw)
    for step in range(1, 4):
        gs = [torch.randn(*s, device="cuda").to(g_dtype) for s in shapes]
        for p, g in zip(ps, gs):
            p.grad = g.to(p.dtype) if p_dtype != torch.float32 else g
        gs_used = [p.grad for p in ps]
        opt.step(div_scale=2.0)
        adam_reference_step(refs, [g.float() for g in gs_used], ms, vs, 1e-2, 0.9, 0.999, 1e-8, 0.05, step, adamw, True,
                            inv_scale=0.5)
        if p_dtype != torch.float32:
            refs = [r.to(p_dtype).float() for r in refs]
    tol = dict(atol=1e-5, rtol=1e-4) if p_dtype == torch.float32 else dict(atol=2e-2, rtol=2e-2)
    for p, r in zip(ps, refs):
        torch.testing.assert_close(p.detach().float(), r, **tol)


@pytest.mark.gpu
def test_multi_tensor_norm_and_scale():
    from colossalai_b200.ops import multi_tensor as mt

    ts = [torch.randn(s, device="cuda", dtype=d) for s, d in
          [(100000, torch.float32), (77, torch.bfloat16), (2048 * 16 * 3 + 1, torch.bfloat16)]]
    tbl = mt.TensorTable(ts, ts)
    tot, per = mt.norm_sq(tbl, "grad", per_tensor=True)
    ref = torch.stack([t.float().pow(2).sum() for t in ts])
    torch.testing.assert_close(per, ref, rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(tot[0], ref.sum(), rtol=1e-3, atol=1e-2)
    outs = [torch.empty_like(t, dtype=torch.float32) for t in ts]
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    mt.scale(mt.TensorTable(ts, outs), 0.25, flag)
    for t, o in zip(ts, outs):
        torch.testing.assert_close(o, t.float() * 0.25)
    assert flag.item() == 0
    ts[1][3] = float("inf")
    mt.scale(mt.TensorTable(ts, outs), 0.25, flag)
    assert flag.item() == 1


@pytest.mark.gpu
def test_mixed_precision_fused_step_matches_reference():
    """bf16 working params + fp32 master through the single-launch fused path vs an fp32 AdamW oracle."""
    from colossalai_b200.amp import MixedPrecisionOptimizer

    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 256, bias=True).cuda()
    ref = torch.nn.Linear(256, 256, bias=True).cuda()
    ref.load_state_dict(lin.state_dict())
    lin = lin.to(torch.bfloat16)
    for p, r in zip(lin.parameters(), ref.parameters()):
        r.data.copy_(p.data.float())
    opt = MixedPrecisionOptimizer(FusedAdam(lin.parameters(), lr=1e-2, weight_decay=0.1), lin, precision="bf16",
                                  max_norm=0.5)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.1)
    for _ in range(3):
        x = torch.randn(64, 256, device="cuda")
        lin(x.bfloat16()).float().pow(2).mean().backward()
        for p, r in zip(lin.parameters(), ref.parameters()):
            r.grad = p.grad.float()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        opt.step()
        opt.zero_grad()
        ropt.step()
        ropt.zero_grad()
        for p, r in zip(lin.parameters(), ref.parameters()):
            master = opt.working_to_master_map[p]
            torch.testing.assert_close(master, r.data, atol=1e-5, rtol=1e-4)
            torch.testing.assert_close(p.data.float(), master.bfloat16().float())
