"""Sharded-vs-single-device oracle for the vision / encoder-decoder / multimodal families (ViT, T5, Whisper, BLIP-2,
SAM) under TP=2 on gloo (reference pattern: tests/test_shardformer/test_model/test_shard_{vit,t5,whisper,blip2,sam}.py)."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.cluster import DeviceMesh
from colossalai_b200.models import build_model
from colossalai_b200.parallel import comm
from colossalai_b200.shardformer import ShardConfig, ShardFormer
from colossalai_b200.tensor.d_tensor import is_distributed_tensor
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _inputs(name):
    g = torch.Generator().manual_seed(11)
    if name == "vit-tiny":
        return dict(pixel_values=torch.randn(2, 3, 32, 32, generator=g), labels=torch.tensor([1, 7]))
    if name == "t5-tiny":
        return dict(input_ids=torch.randint(2, 512, (2, 12), generator=g),
                    attention_mask=torch.tensor([[1] * 12, [1] * 9 + [0] * 3]),
                    labels=torch.randint(2, 512, (2, 8), generator=g))
    if name == "whisper-test":
        return dict(input_features=torch.randn(2, 16, 40, generator=g), labels=torch.randint(3, 512, (2, 8), generator=g))
    if name == "blip2-tiny":
        ids = torch.randint(3, 512, (2, 6), generator=g)
        return dict(pixel_values=torch.randn(2, 3, 32, 32, generator=g), input_ids=ids, labels=ids)
    if name == "sam-tiny":
        return dict(pixel_values=torch.randn(2, 3, 64, 64, generator=g), input_points=torch.rand(2, 3, 2, generator=g) * 64,
                    input_labels=torch.ones(2, 3, dtype=torch.long),
                    labels=(torch.rand(2, 3, 32, 32, generator=g) > 0.5).float())
    raise KeyError(name)


def _run_one(name, parallel_output=True):
    mesh = DeviceMesh(dp=dist.get_world_size() // 2, tp=2)
    torch.manual_seed(1234)
    base = build_model(name)
    sharded = copy.deepcopy(base)
    sc = ShardConfig(tensor_parallel_process_group=mesh.group("tp"), enable_tensor_parallelism=True,
                     parallel_output=parallel_output)
    sharded, _ = ShardFormer(sc).optimize(sharded)
    n_dist = sum(1 for p in sharded.parameters() if is_distributed_tensor(p))
    assert n_dist >= 8, f"{name}: only {n_dist} parameters were sharded"
    inp = _inputs(name)
    out_b, out_s = base(**inp), sharded(**inp)
    torch.testing.assert_close(out_s["loss"], out_b["loss"], atol=3e-5, rtol=1e-4, msg=lambda m: f"loss {name}: {m}")
    out_b["loss"].backward()
    out_s["loss"].backward()
    ref = dict(base.named_parameters())
    checked = 0
    for pname, p in sharded.named_parameters():
        if p.grad is None:
            continue
        g = p.grad
        if is_distributed_tensor(p):
            g = p.gather_fn(g) if hasattr(p, "shard_fn") else comm.all_gather(g, p.dist_shard[0], p.dist_shard[1])
        r = ref[pname].grad
        if g.shape != r.shape:
            g = g[: r.shape[0]]
        torch.testing.assert_close(g, r, atol=5e-5, rtol=2e-3, msg=lambda m: f"{name}.{pname}: {m}")
        checked += 1
    assert checked > 10
    mesh.destroy_mesh_process_groups()


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    for name in ("vit-tiny", "t5-tiny", "whisper-test", "blip2-tiny", "sam-tiny"):
        _run_one(name)
    _run_one("t5-tiny", parallel_output=False)
    # generation on the sharded model reproduces the unsharded tokens
    mesh = DeviceMesh(dp=1, tp=2)
    torch.manual_seed(5)
    base = build_model("t5-tiny").eval()
    sharded, _ = ShardFormer(ShardConfig(tensor_parallel_process_group=mesh.group("tp"),
                                         enable_tensor_parallelism=True)).optimize(copy.deepcopy(base))
    ids = torch.randint(2, 512, (2, 10))
    assert torch.equal(base.generate(ids, max_new_tokens=4), sharded.eval().generate(ids, max_new_tokens=4))
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_shard_encdec_families_tp2():
    spawn(_worker, 2)


def test_vit_pipeline_stage_split():
    """Layer distribution + held-layer bookkeeping of the ViT policy (reference: test_shard_vit.py pipeline branch)."""
    from colossalai_b200.shardformer.policies.vit import ViTForImageClassificationPolicy

    class _SM:   # two-stage manager stub
        num_stages, stage, is_interleave, use_zbv = 2, 1, False, False

        def distribute_layers(self, n): return [n // 2, n - n // 2]
        def get_stage_index(self, per): return (per[0], per[0] + per[1])
        def is_first_stage(self, **k): return False
        def is_last_stage(self, **k): return True

    m = build_model("vit-tiny")
    pol = ViTForImageClassificationPolicy()
    pol.set_model(m)

    class _SC:
        pipeline_stage_manager = _SM()
    pol.shard_config = _SC()
    held = pol.get_held_layers()
    assert m.vit.layers[1] in held and m.vit.layers[0] not in held
    assert m.classifier in held and m.vit.embeddings not in held


if __name__ == "__main__":
    test_shard_encdec_families_tp2()


def _vit_pp_worker(rank, world_size, port):
    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import HybridParallelPlugin
    from colossalai_b200.nn.optimizer import FusedAdam

    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    torch.manual_seed(7)
    base = build_model("vit-tiny")
    model = copy.deepcopy(base)
    ref_opt = torch.optim.AdamW(base.parameters(), lr=1e-2, weight_decay=0.0)
    opt = FusedAdam(model.parameters(), lr=1e-2, weight_decay=0.0)
    booster = Booster(plugin=HybridParallelPlugin(tp_size=1, pp_size=2, precision="fp32", num_microbatches=2))
    model, opt, *_ = booster.boost(model, opt)
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    y = torch.tensor([1, 2, 3, 4])
    out = booster.execute_pipeline(iter([{"pixel_values": x, "labels": y}]), model, lambda o, b: o["loss"], opt,
                                   return_loss=True)
    opt.step()
    total = 0.0
    for i in range(2):
        l = base(pixel_values=x[2 * i: 2 * i + 2], labels=y[2 * i: 2 * i + 2])["loss"] / 2
        l.backward()
        total += l.item()
    ref_opt.step()
    if out["loss"] is not None:
        assert abs(out["loss"].item() - total) < 1e-5
    ref = dict(base.named_parameters())
    for n, p in model.unwrap().named_parameters():
        if p is not None:
            torch.testing.assert_close(p.detach(), ref[n].detach(), atol=1e-5, rtol=1e-4, msg=lambda m: f"vit pp {n}: {m}")
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_vit_pipeline_parallel_matches_single_process():
    spawn(_vit_pp_worker, 2)
