"""GeneralCheckpointIO: unsharded / sharded (bin + safetensors) model and optimizer round trips, async writer
(reference: tests/test_checkpoint_io/test_general_checkpoint_io.py)."""
import os

import pytest
import torch
import torch.nn as nn

from colossalai_b200.checkpoint_io import GeneralCheckpointIO


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 8))


@pytest.mark.parametrize("shard", [False, True])
@pytest.mark.parametrize("safetensors", [False, True])
def test_model_roundtrip(tmp_path, shard, safetensors):
    io = GeneralCheckpointIO()
    m = _model()
    path = str(tmp_path / ("ckpt" if shard else ("model.safetensors" if safetensors else "model.bin")))
    io.save_model(m, path, shard=shard, size_per_shard=0.01, use_safetensors=safetensors)
    if shard:
        files = os.listdir(path)
        assert any(f.endswith(".index.json") for f in files) and len([f for f in files if "0000" in f]) > 1
    m2 = _model()
    with torch.no_grad():
        for p in m2.parameters():
            p.zero_()
    io.load_model(m2, path)
    for a, b in zip(m.parameters(), m2.parameters()):
        torch.testing.assert_close(a, b)


@pytest.mark.parametrize("shard", [False, True])
def test_optimizer_roundtrip(tmp_path, shard):
    io = GeneralCheckpointIO()
    m = _model()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2)
    m(torch.randn(4, 32)).sum().backward()
    opt.step()
    path = str(tmp_path / ("optim_dir" if shard else "optim.bin"))
    io.save_optimizer(opt, path, shard=shard, size_per_shard=0.01)
    m2 = _model()
    opt2 = torch.optim.AdamW(m2.parameters(), lr=123.0)
    io.load_optimizer(opt2, path)
    assert opt2.param_groups[0]["lr"] == pytest.approx(1e-2)
    for (_, s1), (_, s2) in zip(sorted(opt.state_dict()["state"].items()), sorted(opt2.state_dict()["state"].items())):
        torch.testing.assert_close(s1["exp_avg"], s2["exp_avg"])
        torch.testing.assert_close(s1["exp_avg_sq"], s2["exp_avg_sq"])


def test_async_save(tmp_path):
    io = GeneralCheckpointIO()
    m = _model()
    path = str(tmp_path / "async.safetensors")
    io.save_model(m, path, use_safetensors=True, use_async=True)
    io.synchronize()
    m2 = _model()
    with torch.no_grad():
        for p in m2.parameters():
            p.add_(1.0)
    io.load_model(m2, path)
    for a, b in zip(m.parameters(), m2.parameters()):
        torch.testing.assert_close(a, b)
