"""TorchDDPPlugin / TorchFSDPPlugin / LoRA through the Booster API on gloo (reference: tests/test_booster/test_plugin/
test_torch_ddp_plugin.py, test_torch_fsdp_plugin.py, tests/test_lora/test_lora.py)."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.lora import LoraConfig, LoraLinear, apply_lora, merge_lora
from colossalai_b200.booster.plugin import TorchDDPPlugin, TorchFSDPPlugin
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


class _MLP(nn.Module):
    def __init__(self):
        super().__init__()
        self.q_proj = nn.Linear(16, 32)
        self.act = nn.GELU()
        self.o_proj = nn.Linear(32, 4)

    def forward(self, x):
        return self.o_proj(self.act(self.q_proj(x)))


def _ddp_worker(rank, world_size, port, tmp):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    torch.manual_seed(0)
    model = _MLP()
    ref = _MLP()
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    booster = Booster(plugin=TorchDDPPlugin())
    model, opt, *_ = booster.boost(model, opt)
    torch.manual_seed(10)
    full = torch.randn(world_size * 4, 16)
    x = full[rank * 4:(rank + 1) * 4]
    loss = model(x).square().mean()
    booster.backward(loss, opt)
    opt.step()
    ref(full).square().mean().backward()
    ref_opt.step()
    for a, b in zip(model.unwrap().parameters(), ref.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    # no_sync accumulates locally
    with booster.no_sync(model, opt):
        model(x).sum().backward()
    # checkpoint round trip (rank 0 writes)
    path = os.path.join(tmp, "ddp_model.pt")
    booster.save_model(model, path)
    dist.barrier()
    with torch.no_grad():
        for p in model.unwrap().parameters():
            p.add_(1.0)
    booster.load_model(model, path)
    for a, b in zip(model.unwrap().parameters(), ref.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    booster.save_model(model, os.path.join(tmp, "ddp_sharded"), shard=True, size_per_shard=1)
    dist.barrier()
    booster.load_model(model, os.path.join(tmp, "ddp_sharded"))
    dist.barrier()
    dist.destroy_process_group()


def _fsdp_worker(rank, world_size, port, tmp):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="nccl", verbose=False)
    torch.manual_seed(0)
    model = _MLP()
    ref = _MLP()
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    booster = Booster(plugin=TorchFSDPPlugin())
    model, opt, *_ = booster.boost(model, opt)
    torch.manual_seed(10)
    full = torch.randn(world_size * 4, 16).cuda()
    ref = ref.cuda()
    x = full[rank * 4:(rank + 1) * 4]
    loss = model(x).square().mean()
    booster.backward(loss, opt)
    opt.step()
    ref(full).square().mean().backward()
    ref_opt.step()
    path = os.path.join(tmp, "fsdp_model.pt")
    booster.save_model(model, path)
    dist.barrier()
    if rank == 0:
        sd = torch.load(path, weights_only=True)
        for k, v in ref.state_dict().items():
            torch.testing.assert_close(sd[k], v.cpu(), rtol=1e-5, atol=1e-6)
    booster.load_model(model, path)
    booster.save_optimizer(opt, os.path.join(tmp, "fsdp_optim.pt"))
    dist.barrier()
    booster.load_optimizer(opt, os.path.join(tmp, "fsdp_optim.pt"))
    dist.barrier()
    dist.destroy_process_group()


@rerun_if_address_is_in_use()
def test_torch_ddp_plugin():
    with tempfile.TemporaryDirectory() as tmp:
        spawn(_ddp_worker, 2, tmp=tmp)


@pytest.mark.gpu
@rerun_if_address_is_in_use()
def test_torch_fsdp_plugin():
    if torch.cuda.device_count() < 2:
        pytest.skip("FSDP needs >= 2 accelerator devices")
    with tempfile.TemporaryDirectory() as tmp:
        spawn(_fsdp_worker, 2, tmp=tmp)


def test_lora_apply_train_save_load_merge():
    torch.manual_seed(0)
    model = _MLP()
    base = {k: v.clone() for k, v in model.state_dict().items()}
    model = apply_lora(model, LoraConfig(r=4, lora_alpha=8, target_modules=["q_proj", "o_proj"]))
    assert isinstance(model.q_proj, LoraLinear)
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable and all("lora_" in n for n in trainable)
    x = torch.randn(8, 16)
    y0 = model(x)
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.5)
    for _ in range(3):
        opt.zero_grad()
        model(x).square().mean().backward()
        opt.step()
    assert not torch.allclose(model(x), y0)
    torch.testing.assert_close(model.q_proj.base_layer.weight, base["q_proj.weight"])      # base frozen
    with tempfile.TemporaryDirectory() as tmp:
        from colossalai_b200.booster.lora import save_lora_adapters

        save_lora_adapters(model, tmp)
        assert os.path.exists(os.path.join(tmp, "adapter_config.json"))
        torch.manual_seed(0)
        fresh = apply_lora(_MLP(), None, pretrained_dir=tmp)
        torch.testing.assert_close(fresh(x), model(x))
    y = model(x)
    merge_lora(model)
    torch.testing.assert_close(model(x), y, rtol=1e-4, atol=1e-5)
