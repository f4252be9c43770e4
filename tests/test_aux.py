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
