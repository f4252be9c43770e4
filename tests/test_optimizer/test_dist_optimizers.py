"""TP-aware optimizers: Lamb / CAME / Adafactor (/ GaLore) applied to tensor-parallel shards through the hybrid plugin
must update the weights exactly like the plain optimizer does on the full weights — their trust ratios, factored
second moments and RMS clipping are statistics of the WHOLE parameter (reference: tests/test_optimizer/
test_dist_{lamb,came,adafactor,galore}.py)."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.plugin import HybridParallelPlugin
from colossalai_b200.models import build_model
from colossalai_b200.nn.optimizer import CAME, Adafactor, Lamb
from colossalai_b200.parallel import comm
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _gather(p):
    if hasattr(p, "gather_fn"):
        return p.gather_fn(p)
    if hasattr(p, "dist_shard"):
        return comm.all_gather(p.detach(), p.dist_shard[0], p.dist_shard[1])
    return p.detach()


def _run(make_opt, tag, plugin_kw, atol, mean_only=False, expect_dist=True):
    torch.manual_seed(42)
    base = build_model("llama-tiny")
    model = copy.deepcopy(base)
    ref_opt = make_opt(base.parameters())
    opt = make_opt(model.parameters())
    plugin = HybridParallelPlugin(precision="fp32", **plugin_kw)
    booster = Booster(plugin=plugin)
    model, opt, *_ = booster.boost(model, opt)
    assert not expect_dist or "Dist" in type(opt.optim).__name__, f"{tag}: expected the distributed variant, got {type(opt.optim).__name__}"
    dp_rank = plugin.pg_mesh.axis_rank("dp")
    torch.manual_seed(100)
    ids = torch.randint(0, 512, (2 * plugin.dp_size, 32))
    for _ in range(3):
        mine = ids[2 * dp_rank: 2 * dp_rank + 2]
        booster.backward(model(input_ids=mine, labels=mine)["loss"], opt)
        opt.step()
        opt.zero_grad()
        base(input_ids=ids, labels=ids)["loss"].backward()
        ref_opt.step()
        ref_opt.zero_grad()
    ref = dict(base.named_parameters())
    worst = 0.0
    for name, p in model.unwrap().named_parameters():
        full = _gather(p)
        r = ref[name]
        if full.shape != r.shape:
            full = full[: r.shape[0]]
        worst = max(worst, (full - r.detach()).abs().max().item())
        if mean_only:       # SVD-based projections amplify 1e-8 input noise on single elements: compare in the mean
            assert (full - r.detach()).abs().mean().item() <= atol, f"{tag} {name}"
        else:
            torch.testing.assert_close(full, r.detach(), atol=atol, rtol=5e-3, msg=lambda m: f"{tag} {name}: {m}")
    del plugin
    return worst


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    cases = [
        ("lamb", lambda ps: Lamb(ps, lr=1e-2, weight_decay=0.01)),
        ("came", lambda ps: CAME(ps, lr=1e-3, betas=(0.9, 0.999, 0.9999), weight_decay=0.0)),
        ("adafactor", lambda ps: Adafactor(ps, lr=1e-3, relative_step=False, scale_parameter=True, weight_decay=0.0)),
    ]
    for tag, mk in cases:
        _run(mk, f"{tag}/tp2", dict(tp_size=2, pp_size=1), atol=5e-5)
    # user-defined param groups in NON-model order (decay for matrices, none + higher lr for norms): membership must
    # follow the parameters through sharding by name, not by position
    def two_groups(ps):
        ps = list(ps)
        return torch.optim.AdamW([dict(params=[p for p in ps if p.dim() >= 2], weight_decay=0.1),
                                  dict(params=[p for p in ps if p.dim() < 2], weight_decay=0.0, lr=5e-2)], lr=1e-2)

    _run(two_groups, "adamw-two-groups/tp2", dict(tp_size=2, pp_size=1), atol=2e-4, expect_dist=False)
    # GaLore projects the FULL gradient: shards are gathered (block-aware for fused q|k|v), projected, updated, re-split
    from colossalai_b200.nn.optimizer import GaLoreAdamW8bit

    def galore(ps):
        ps = list(ps)
        mats = [p for p in ps if p.dim() == 2]
        rest = [p for p in ps if p.dim() != 2]
        return GaLoreAdamW8bit([dict(params=mats, rank=8, update_proj_gap=10, scale=0.25, proj_type="std"),
                                dict(params=rest)], lr=1e-3, weight_decay=0.0, nbits=32)

    # the SVD basis is chaotic w.r.t. 1e-8 input noise (near-degenerate singular values) and Adam's element-wise
    # normalisation is not rotation invariant, so pin the basis: everything else must then agree exactly
    from colossalai_b200.nn.optimizer.galore import GaLoreProjector

    def fixed_basis(self, g, side):
        n = g.shape[0] if side == "left" else g.shape[1]
        q, _ = torch.linalg.qr(torch.randn(n, self.rank, generator=torch.Generator().manual_seed(n)))
        return q if side == "left" else q.t()

    orig = GaLoreProjector._svd
    GaLoreProjector._svd = fixed_basis
    try:
        _run(galore, "galore/tp2", dict(tp_size=2, pp_size=1), atol=2e-5, mean_only=True)   # eps-level Adam flips on ~0 coords
    finally:
        GaLoreProjector._svd = orig
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_distributed_optimizers_match_single_process_tp2():
    spawn(_worker, 2)


def test_lars_and_lamb_formulas():
    """Single-tensor sanity of the trust-ratio optimizers against hand-computed first steps."""
    from colossalai_b200.nn.optimizer import Lars

    w = torch.nn.Parameter(torch.tensor([3.0, 4.0]))
    w.grad = torch.tensor([0.6, 0.8])
    opt = Lamb([w], lr=0.1, betas=(0.9, 0.999), eps=0.0, weight_decay=0.0)
    opt.step()
    # first step: m = 0.1 g, v = 0.001 g^2 -> update = m / sqrt(v) = 0.1/sqrt(0.001) * sign(g) (per element), trust = |w|/|u|
    u = torch.full((2,), 0.1 / 0.001 ** 0.5)
    expect = torch.tensor([3.0, 4.0]) - 0.1 * (5.0 / u.norm()) * u
    torch.testing.assert_close(w.detach(), expect, atol=1e-5, rtol=1e-5)
    w2 = torch.nn.Parameter(torch.tensor([3.0, 4.0]))
    w2.grad = torch.tensor([0.6, 0.8])
    Lars([w2], lr=0.1, momentum=0.0, weight_decay=0.0).step()
    assert torch.isfinite(w2).all() and not torch.equal(w2.detach(), torch.tensor([3.0, 4.0]))
