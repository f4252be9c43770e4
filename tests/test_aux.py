"""Aux subsystems: watchdog, training-state resume, comm profiler, performance evaluator."""
import time

import torch
import torch.distributed as dist
from torch.utils.data import TensorDataset

from colossalai_b200.checkpoint_io.training_state import (StatefulDistributedSampler, load_training_state,
                                                          save_training_state)
from colossalai_b200.utils.profiler import CommProfiler, PerformanceEvaluator, get_profile_context
from colossalai_b200.utils.watchdog import StepWatchdog


def test_watchdog_fires_and_beats(tmp_path):
    hits = []
    wd = StepWatchdog(timeout_s=0.3, on_timeout=hits.append, abort=False, poll_s=0.05,
                      heartbeat_file=str(tmp_path / "hb")).start()
    for _ in range(3):
        time.sleep(0.1)
        wd.beat()
    assert not wd.fired and (tmp_path / "hb").read_text().startswith("3 ")
    time.sleep(0.6)
    wd.stop()
    assert wd.fired and hits and hits[0] > 0.3


def test_training_state_resume(tmp_path):
    ds = TensorDataset(torch.arange(20))
    s = StatefulDistributedSampler(ds, num_replicas=2, rank=0, shuffle=True, seed=1)
    s.set_epoch(3)
    full = list(iter(s))
    torch.manual_seed(5)
    torch.rand(3)
    save_training_state(str(tmp_path), epoch=3, step=17, sample_start_index=4)
    expect = torch.rand(4)
    s2 = StatefulDistributedSampler(ds, num_replicas=2, rank=0, shuffle=True, seed=1)
    torch.manual_seed(999)
    st = load_training_state(str(tmp_path), s2)
    assert st["step"] == 17 and list(iter(s2)) == full[4:] and len(s2) == len(full) - 4
    torch.testing.assert_close(torch.rand(4), expect)


def test_profilers_cpu():
    with get_profile_context(False, 1, 1) as prof:
        prof.step()
    ev = PerformanceEvaluator(model_numel=1000, num_layers=2, hidden_size=8, vocab_size=32, ignore_steps=1)
    for step in range(3):
        ev.on_step_start(step)
        time.sleep(0.01)
        ev.on_step_end(torch.zeros(2, 16, dtype=torch.long))
    out = ev.on_fit_end()
    assert ev.num_samples == 4 and out["samples_per_sec"] > 0
    prof = CommProfiler()
    with prof:
        assert dist.all_reduce.__name__ == "wrapped"
    assert dist.all_reduce.__name__ != "wrapped" and "collective" in prof.result_str()


def test_async_safetensors_seed_and_decode_workspace(tmp_path):
    """utils.safetensors (reference utils/safetensors.py), testing.random.seed_all, FDIntermTensors workspace."""
    import torch

    from colossalai_b200.inference.flash_decoding_utils import FDIntermTensors
    from colossalai_b200.testing.random import seed_all
    from colossalai_b200.utils.safetensors import (create_pinned_state_dict, load_flat, move_and_save, save_nested)

    seed_all(123)
    a = torch.rand(3)
    seed_all(123)
    assert torch.equal(a, torch.rand(3))
    sd = {"w": torch.randn(5, 7), "b": torch.arange(4, dtype=torch.int64), "h": torch.randn(3).bfloat16()}
    pinned = create_pinned_state_dict(sd)
    w = move_and_save(str(tmp_path / "m.safetensors"), sd, pinned)
    w.synchronize()
    from safetensors.torch import load_file

    back = load_file(str(tmp_path / "m.safetensors"))
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    osd = {"state": {0: {"step": torch.tensor(3.0), "exp_avg": torch.randn(4), "flag": 7},
                     1: {"step": torch.tensor(3.0), "exp_avg": torch.randn(2, 2)}},
           "param_groups": [{"lr": 1e-3, "betas": [0.9, 0.95], "params": [0, 1]}]}
    w = save_nested(str(tmp_path / "o.safetensors"), osd)
    w.synchronize()
    got = load_flat(str(tmp_path / "o.safetensors"))
    assert got["param_groups"] == osd["param_groups"] and got["state"][0]["flag"] == 7
    assert torch.equal(got["state"][1]["exp_avg"], osd["state"][1]["exp_avg"])
    fd = FDIntermTensors()
    fd._reset()
    assert fd.views(2, 4, 3, 8) is None
    fd.initialize(4, 4, 3, 8, device="cpu")
    o, ml = fd.views(2, 4, 3, 8)
    assert o.shape == (2, 4, 3, 8) and ml.shape == (2, 4, 3, 2) and o.data_ptr() == fd.mid_output.data_ptr()
    assert fd.views(8, 4, 3, 8) is None and fd.exp_sums.shape == (4, 4, 3)
    fd._reset()
