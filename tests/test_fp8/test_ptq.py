"""Post-training quantisation: GPTQ (beats round-to-nearest, packs losslessly) and SmoothQuant W8A8 (smoothing keeps
the float function, W8A8 model stays close) — reference: tests/test_legacy/test_infer_ops + quant examples."""
import copy

import torch
import torch.nn as nn

from colossalai_b200.models import build_model
from colossalai_b200.quantization.gptq import GPTQ, QuantLinear, Quantizer, gptq_quantize_model, pack_rows, unpack_rows
from colossalai_b200.quantization.smoothquant import (W8A8Linear, get_act_scales, smooth_and_quantize_model,
                                                      smooth_ln_fcs)


def test_pack_roundtrip():
    for bits in (2, 4, 8):
        q = torch.randint(0, 2 ** bits, (64, 6))
        assert torch.equal(unpack_rows(pack_rows(q, bits), bits), q)


def test_gptq_beats_rtn_and_packs():
    torch.manual_seed(0)
    lin = nn.Linear(128, 64)
    X = torch.randn(2048, 16) @ torch.randn(16, 128) + 0.1 * torch.randn(2048, 128)     # correlated inputs
    ref = lin(X)
    W, qz = lin.weight.data.clone(), Quantizer(4)
    rtn = torch.zeros_like(W)
    for g in range(0, 128, 32):
        qz.find_params(W[:, g:g + 32])
        rtn[:, g:g + 32] = qz.quantize(W[:, g:g + 32])
    err_rtn = (nn.functional.linear(X, rtn, lin.bias) - ref).pow(2).mean()
    for actorder in (False, True):
        l2 = copy.deepcopy(lin)
        solver = GPTQ(l2)
        solver.add_batch(X[:1024])
        solver.add_batch(X[1024:])
        scales, zeros, g_idx = solver.fasterquant(bits=4, group_size=32, actorder=actorder)
        err = (l2(X) - ref).pow(2).mean()
        assert err < 0.3 * err_rtn, (err, err_rtn)
        ql = QuantLinear(4, 32, 128, 64)
        ql.pack(l2, scales, zeros, g_idx)
        torch.testing.assert_close(ql(X).float(), l2(X), atol=2e-2, rtol=2e-2)
        assert ql.qweight.shape == (128 * 4 // 32, 64) and ql.qweight.dtype == torch.int32


def test_gptq_model_level():
    torch.manual_seed(0)
    model = build_model("llama-tiny").float().eval()
    batches = [dict(input_ids=torch.randint(0, 512, (2, 32))) for _ in range(4)]
    ids = torch.randint(0, 512, (2, 16))
    ref = model(input_ids=ids)["logits"]
    gptq_quantize_model(model, batches, bits=8, group_size=32)
    assert sum(isinstance(m, QuantLinear) for m in model.modules()) == 4 * 4
    got = model(input_ids=ids)["logits"]
    assert torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0) > 0.99


def test_smoothquant():
    torch.manual_seed(0)
    model = build_model("llama-tiny").float().eval()
    ids = torch.randint(0, 512, (2, 16))
    ref = model(input_ids=ids)["logits"]
    batches = [dict(input_ids=torch.randint(0, 512, (2, 32))) for _ in range(2)]
    scales = get_act_scales(model, batches)
    layer = model.model.layers[0]
    s = smooth_ln_fcs(layer.input_layernorm, [layer.self_attn.qkv_proj], scales["model.layers.0.self_attn.qkv_proj"])
    assert s.shape == (64,)
    torch.testing.assert_close(model(input_ids=ids)["logits"], ref, atol=1e-4, rtol=1e-4)   # float function unchanged
    smooth_and_quantize_model(model, batches)
    assert sum(isinstance(m, W8A8Linear) for m in model.modules()) == 4 * 4
    got = model(input_ids=ids)["logits"]
    assert torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0) > 0.99
    lin = nn.Linear(32, 16)
    x = torch.randn(4, 7, 32)
    torch.testing.assert_close(W8A8Linear.from_float(lin)(x), lin(x), atol=0.05, rtol=0.05)
