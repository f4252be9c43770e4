"""LoRA end to end through the Booster: `enable_lora` -> plugin boost (ZeRO-2 bf16, DDP) -> only the adapters train,
the adapters save / reload, merge reproduces the adapted model (reference: tests/test_lora/test_lora.py)."""
import os

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.lora import LoraConfig, LoraLinear, merge_lora
from colossalai_b200.booster.plugin import LowLevelZeroPlugin, TorchDDPPlugin
from colossalai_b200.models import build_model
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _run(plugin, tmp, tag):
    torch.manual_seed(0)
    model = build_model("llama-tiny")
    base = {k: v.detach().clone() for k, v in model.state_dict().items()}
    booster = Booster(plugin=plugin)
    model = booster.enable_lora(model, lora_config=LoraConfig(r=4, lora_alpha=8, target_modules=["qkv_proj", "o_proj"]))
    n_lora = sum(isinstance(m, LoraLinear) for m in model.modules())
    assert n_lora == 2 * model.cfg.num_hidden_layers
    params = [p for p in model.parameters() if p.requires_grad]
    assert params and len(params) == 2 * n_lora
    opt = torch.optim.AdamW(params, lr=2e-2)
    model, opt, *_ = booster.boost(model, opt)
    ids = (torch.arange(32)[None] * 3 + torch.tensor([[1], [5]])) % 512
    losses = []
    for _ in range(8):
        loss = model(input_ids=ids, labels=ids)["loss"]
        booster.backward(loss, opt)
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
    assert losses[-1] < losses[0] - 0.05, (tag, losses)
    inner = model.unwrap() if hasattr(model, "unwrap") else model
    inner = inner.module if hasattr(inner, "module") and not hasattr(inner, "cfg") else inner
    sd = inner.state_dict()
    for k, v in base.items():                    # frozen base weights did not move
        key = k if k in sd else k.replace(".weight", ".base_layer.weight")
        torch.testing.assert_close(sd[key].float(), v.float(), atol=1e-2 if "bf16" in tag else 0, rtol=0,
                                   msg=lambda m: f"{tag} {k}: {m}")
    path = os.path.join(tmp, tag)
    booster.save_lora_as_pretrained(model, path)
    dist.barrier()
    assert os.path.exists(os.path.join(path, "adapter_config.json"))
    # reload the adapters into a fresh base model: same function
    torch.manual_seed(0)
    fresh = booster.enable_lora(build_model("llama-tiny"), pretrained_dir=path)
    probe = torch.randint(0, 512, (1, 16), generator=torch.Generator().manual_seed(9))
    inner.eval()
    with torch.no_grad():
        ref = inner(input_ids=probe)["logits"].float()
        got = fresh.to(ref.device).eval()(input_ids=probe)["logits"].float()
    torch.testing.assert_close(got, ref, atol=6e-2 if "bf16" in tag else 1e-4, rtol=5e-2)
    merged = merge_lora(fresh, unload=True)
    with torch.no_grad():
        torch.testing.assert_close(merged(input_ids=probe)["logits"].float(), got, atol=1e-3, rtol=1e-3)
    assert not any(isinstance(m, LoraLinear) for m in merged.modules())


def _worker(rank, world_size, port, tmp):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    _run(TorchDDPPlugin(), tmp, "ddp")
    _run(LowLevelZeroPlugin(stage=2, precision="bf16"), tmp, "zero2_bf16")
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_lora_through_booster_plugins(tmp_path):
    spawn(_worker, 2, tmp=str(tmp_path))
