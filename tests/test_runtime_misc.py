"""Small runtime pieces without a test of their own: accelerator API, Config, DistributedLogger, Model / Optimizer wrappers
(reference: tests/test_config, tests/test_booster/test_accelerator.py, colossalai/interface/*)."""
import logging

import pytest
import torch
import torch.nn as nn

from colossalai_b200.accelerator import CpuAccelerator, get_accelerator, set_accelerator
from colossalai_b200.context import Config
from colossalai_b200.interface import ModelWrapper, OptimizerWrapper
from colossalai_b200.logging import disable_existing_loggers, get_dist_logger


def test_cpu_accelerator_api():
    set_accelerator("cpu")
    acc = get_accelerator()
    assert isinstance(acc, CpuAccelerator) and acc.name == "cpu" and acc.communication_backend == "gloo"
    assert acc.get_current_device() == torch.device("cpu")
    acc.manual_seed(3)
    a = torch.rand(4)
    state = acc.get_rng_state()
    b = torch.rand(4)
    acc.set_rng_state(state)
    torch.testing.assert_close(torch.rand(4), b)                      # the state round-trips
    acc.manual_seed(3)
    torch.testing.assert_close(torch.rand(4), a)
    acc.synchronize()
    acc.empty_cache()
    assert acc.memory_allocated() >= 0 and acc.max_memory_allocated() >= 0
    free, total = acc.mem_get_info()
    assert 0 < free <= total
    with acc.autocast(enabled=True, dtype=torch.bfloat16):
        y = torch.nn.functional.linear(torch.randn(2, 8), torch.randn(4, 8))
    assert y.dtype == torch.bfloat16
    with acc.stream(acc.Stream()):                                    # no-op objects on the CPU tier
        acc.Event().record()
    from colossalai_b200.accelerator import auto_set_accelerator

    auto_set_accelerator()                                            # back to what the box has


def test_config_attribute_access_and_files(tmp_path):
    cfg = Config({"parallel": {"tensor": {"size": 2, "mode": "1d"}}, "fp16": {"mode": None}, "lr": 1e-3})
    assert cfg.parallel.tensor.size == 2 and cfg["parallel"]["tensor"]["mode"] == "1d" and cfg.lr == 1e-3
    cfg.update({"lr": 2e-3}, clip=1.0)
    assert cfg.lr == 2e-3 and cfg.clip == 1.0
    cfg.new_section = {"a": 1}
    assert cfg.new_section.a == 1                                     # nested dicts become Configs
    with pytest.raises((AttributeError, KeyError)):
        _ = cfg.missing_key
    f = tmp_path / "conf.py"
    f.write_text("BATCH = 8\nmodel = dict(hidden=64, layers=[1, 2])\n")
    loaded = Config.from_file(f)
    assert loaded.BATCH == 8 and loaded.model.hidden == 64 and loaded.model.layers == [1, 2]


def test_dist_logger_levels_and_file(tmp_path, capsys):
    log = get_dist_logger("cb200_test_logger")
    assert get_dist_logger("cb200_test_logger") is log                # one instance per name
    log.set_level("WARNING")
    with pytest.raises(Exception):
        log.set_level("LOUD")
    log.log_to_file(tmp_path, mode="w", level="INFO")
    log.set_level("INFO")
    log.info("hello from rank 0", ranks=[0])
    log.info("not for this rank", ranks=[7])
    log.warning("careful")
    for h in logging.getLogger("cb200_test_logger").handlers:
        h.flush()
    text = "".join(p.read_text() for p in tmp_path.iterdir())
    assert "hello from rank 0" in text and "careful" in text and "not for this rank" not in text
    disable_existing_loggers(include=["cb200_test_logger"])


def test_model_and_optimizer_wrappers():
    net = nn.Linear(4, 2)
    wrapped = ModelWrapper(net)
    assert wrapped.unwrap() is net and wrapped.in_features == 4       # attribute fall-through
    x = torch.randn(3, 4)
    torch.testing.assert_close(wrapped(x), net(x))
    opt = OptimizerWrapper(torch.optim.SGD(net.parameters(), lr=0.1))
    assert opt.param_groups[0]["lr"] == 0.1 and opt.unwrap().__class__.__name__ == "SGD"
    loss = wrapped(x).square().sum()
    opt.backward(loss)
    norm = opt.get_grad_norm()
    assert norm is None or norm >= 0
    opt.clip_grad_by_norm(0.5)
    total = torch.sqrt(sum(p.grad.square().sum() for p in net.parameters()))
    assert total <= 0.5 + 1e-5
    opt.clip_grad_by_value(0.01)
    assert max(p.grad.abs().max() for p in net.parameters()) <= 0.01 + 1e-8
    before = [p.detach().clone() for p in net.parameters()]
    opt.step()
    opt.zero_grad()
    assert any(not torch.equal(a, b) for a, b in zip(before, net.parameters()))
    sd = opt.state_dict()
    opt.load_state_dict(sd)
    # backward_by_grad: gradient of an intermediate tensor
    y = wrapped(x)
    opt.backward_by_grad(y, torch.ones_like(y))
    assert net.weight.grad is not None


def test_group_key_is_stable_and_not_an_address():
    """Workspace caches are keyed by torch's unique group name, not by `id(group)` (which a new group can reuse)."""
    import torch.distributed as dist

    from colossalai_b200.parallel import comm
    from colossalai_b200.testing import free_port

    assert comm.group_key(None)[0] == "id"                  # no process group yet: falls back, nothing to alias
    dist.init_process_group("gloo", rank=0, world_size=1, init_method=f"tcp://127.0.0.1:{free_port()}")
    try:
        world = comm.group_key(None)
        assert world == comm.group_key(dist.group.WORLD) and world[0] == "pg"
        g1 = dist.new_group([0])
        k1 = comm.group_key(g1)
        assert k1[0] == "pg" and k1 != world
        dist.destroy_process_group(g1)
        del g1
        g2 = dist.new_group([0])
        assert comm.group_key(g2) != k1                       # a new group never inherits the old one's key
    finally:
        dist.destroy_process_group()
