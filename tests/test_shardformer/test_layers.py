"""Layer-level oracles of the parallel building blocks on 2 gloo ranks: every sharded layer is built from a native
module with `from_native_module` and compared with it - output, input gradient and (gathered) parameter gradients.
Reference: tests/test_shardformer/test_layer/test_{linear_1d,qkv_fused_linear_1d,gpt2_qkv_fused_linear_1d,embedding,
vocab_parallel_embedding_1d,layernorm,dropout,dist_crossentropy,sequence_parallel}.py."""
import copy

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

import colossalai_b200
from colossalai_b200.parallel import comm
from colossalai_b200.shardformer.layer import (DropoutForParallelInput, DropoutForReplicatedInput, Embedding1D,
                                                FusedLayerNorm, FusedLinear1D_Col, FusedRMSNorm,
                                                GPT2FusedLinearConv1D_Col, GPT2FusedLinearConv1D_Row, Linear1D_Col,
                                                Linear1D_Row, PaddingEmbedding, PaddingLMHead,
                                                VocabParallelEmbedding1D, VocabParallelLMHead1D, cross_entropy_1d,
                                                dist_log_prob_1d)
from colossalai_b200.shardformer.layer.utils import SeqParallelUtils
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn

TOL = dict(atol=1e-5, rtol=1e-5)


def _full_grad(p):
    g = p.grad.detach()
    if hasattr(p, "dist_shard"):
        dim, group = p.dist_shard
        return comm.all_gather(g, dim, group)
    if hasattr(p, "gather_fn"):
        return p.gather_fn(g)
    return g


def _shard(t, dim):
    return t.chunk(dist.get_world_size(), dim=dim)[dist.get_rank()].contiguous()


def _check_linear_col_row():
    torch.manual_seed(0)
    fc1, fc2 = nn.Linear(16, 32), nn.Linear(32, 16)
    col = Linear1D_Col.from_native_module(copy.deepcopy(fc1), process_group=None)
    row = Linear1D_Row.from_native_module(copy.deepcopy(fc2), process_group=None)
    assert col.weight.shape == (16, 16) and col.bias.shape == (16,) and row.weight.shape == (16, 16)
    x = torch.randn(3, 5, 16)
    xr, xs = x.clone().requires_grad_(), x.clone().requires_grad_()
    ref = fc2(F.gelu(fc1(xr)))
    out = row(F.gelu(col(xs)))                       # the Megatron pair: column -> elementwise -> row
    torch.testing.assert_close(out, ref, **TOL)
    ref.square().sum().backward()
    out.square().sum().backward()
    torch.testing.assert_close(xs.grad, xr.grad, **TOL)
    torch.testing.assert_close(_full_grad(col.weight), fc1.weight.grad, **TOL)
    torch.testing.assert_close(_full_grad(col.bias), fc1.bias.grad, **TOL)
    torch.testing.assert_close(_full_grad(row.weight), fc2.weight.grad, **TOL)
    torch.testing.assert_close(row.bias.grad, fc2.bias.grad, **TOL)
    # gather_output: the column linear alone equals the native layer
    g = Linear1D_Col.from_native_module(copy.deepcopy(fc1), process_group=None, gather_output=True)
    torch.testing.assert_close(g(x), fc1(x), **TOL)
    # parallel_input=False: the row linear splits a replicated input itself
    r = Linear1D_Row.from_native_module(copy.deepcopy(fc2), process_group=None, parallel_input=False)
    h = torch.randn(4, 32)
    torch.testing.assert_close(r(h), fc2(h), **TOL)
    # skip_bias_add hands the bias back un-added (fused into the next kernel by the caller)
    s = Linear1D_Col.from_native_module(copy.deepcopy(fc1), process_group=None, skip_bias_add=True)
    o, b = s(x)
    torch.testing.assert_close(comm.all_gather(o + b, -1, None), fc1(x), **TOL)
    with pytest.raises(ValueError):
        Linear1D_Col.from_native_module(nn.Linear(8, 7), process_group=None)        # 7 columns over 2 ranks


def _check_sequence_parallel_linears():
    """Megatron SP pair: the column linear gathers the sequence shard, the row linear reduce-scatters - both orders of
    the sequence dim ([S, B, H] and [B, S, H]), ring variant included; the row bias is added after the scatter, so its
    gradient is partial until `allreduce_partial_data_grad`."""
    torch.manual_seed(1)
    fc1, fc2 = nn.Linear(16, 32), nn.Linear(32, 16)
    for mode in ("split_gather", "ring"):
        for dim, shape in ((0, (8, 2, 16)), (1, (2, 8, 16))):
            col = Linear1D_Col.from_native_module(copy.deepcopy(fc1), process_group=None, seq_parallel_mode=mode,
                                                  seq_parallel_dim=dim)
            row = Linear1D_Row.from_native_module(copy.deepcopy(fc2), process_group=None, seq_parallel_mode=mode,
                                                  seq_parallel_dim=dim)
            ref1, ref2 = copy.deepcopy(fc1), copy.deepcopy(fc2)
            x = torch.randn(*shape)
            xr = x.clone().requires_grad_()
            xs = _shard(x, dim).requires_grad_()
            ref = ref2(F.gelu(ref1(xr)))
            out = row(F.gelu(col(xs)))
            assert out.shape[dim] == shape[dim] // 2
            torch.testing.assert_close(out, _shard(ref, dim), **TOL)
            ref.square().sum().backward()
            out.square().sum().backward()
            torch.testing.assert_close(xs.grad, _shard(xr.grad, dim), **TOL)
            torch.testing.assert_close(_full_grad(col.weight), ref1.weight.grad, **TOL)
            torch.testing.assert_close(_full_grad(row.weight), ref2.weight.grad, **TOL)
            assert SeqParallelUtils.is_sp_partial_derived_param(row.bias)
            holder = nn.ParameterList([row.bias])
            SeqParallelUtils.allreduce_partial_data_grad(process_group=None, model=holder)
            torch.testing.assert_close(row.bias.grad, ref2.bias.grad, **TOL)
    # pre_gathered: the caller gathered once for several column linears, dX stays partial for its reduce-scatter
    from colossalai_b200.shardformer.layer._operation import gather_forward_reducescatter_backward

    q, k = nn.Linear(16, 16), nn.Linear(16, 8)
    cq = Linear1D_Col.from_native_module(copy.deepcopy(q), process_group=None, seq_parallel_mode="pre_gathered")
    ck = Linear1D_Col.from_native_module(copy.deepcopy(k), process_group=None, seq_parallel_mode="pre_gathered")
    x = torch.randn(2, 8, 16)
    xr = x.clone().requires_grad_()
    xs = _shard(x, 1).requires_grad_()
    full = gather_forward_reducescatter_backward(xs, None, 1)
    (cq(full).square().sum() + ck(full).sum()).backward()
    (q(xr).square().sum() + k(xr).sum()).backward()
    torch.testing.assert_close(xs.grad, _shard(xr.grad, 1), **TOL)
    torch.testing.assert_close(_full_grad(cq.weight), q.weight.grad, **TOL)


def _check_fused_linears():
    torch.manual_seed(2)
    # [q | k | v] with unequal blocks (GQA): every block is split over the ranks separately
    fused = nn.Linear(16, 16 + 8 + 8)
    col = FusedLinear1D_Col.from_native_module(copy.deepcopy(fused), process_group=None, split_sizes=[16, 8, 8])
    assert col.weight.shape == (16, 16) and col.local_split_sizes == [8, 4, 4]
    x = torch.randn(4, 16)
    q, k, v = col(x).split(col.local_split_sizes, -1)
    rq, rk, rv = fused(x).split([16, 8, 8], -1)
    torch.testing.assert_close(q, _shard(rq, -1), **TOL)
    torch.testing.assert_close(k, _shard(rk, -1), **TOL)
    torch.testing.assert_close(v, _shard(rv, -1), **TOL)
    col(x).sum().backward()
    fused(x).sum().backward()
    torch.testing.assert_close(_full_grad(col.weight), fused.weight.grad, **TOL)
    torch.testing.assert_close(_full_grad(col.bias), fused.bias.grad, **TOL)
    # GPT-2 Conv1D layout ([in, out] weights): c_attn column + c_proj row vs the native pair
    import transformers

    c_attn, c_proj = transformers.pytorch_utils.Conv1D(48, 16), transformers.pytorch_utils.Conv1D(16, 16)
    with torch.no_grad():
        c_attn.bias.normal_()
        c_proj.bias.normal_()
    pc = GPT2FusedLinearConv1D_Col.from_native_module(copy.deepcopy(c_attn), process_group=None, split_sizes=[16, 16, 16])
    pr = GPT2FusedLinearConv1D_Row.from_native_module(copy.deepcopy(c_proj), process_group=None)
    xr, xs = x.clone().requires_grad_(), x.clone().requires_grad_()
    rq, rk, rv = c_attn(xr).split(16, -1)
    sq, sk, sv = pc(xs).split(8, -1)
    ref = c_proj(rq * rk + rv)
    out = pr(sq * sk + sv)
    torch.testing.assert_close(out, ref, **TOL)
    ref.sum().backward()
    out.sum().backward()
    torch.testing.assert_close(xs.grad, xr.grad, **TOL)
    # (the parallel layers keep the weight in linear layout [out, in]: Conv1D's [in, out] transposed once at conversion)
    torch.testing.assert_close(_full_grad(pc.weight), c_attn.weight.grad.t(), **TOL)
    torch.testing.assert_close(_full_grad(pr.weight), c_proj.weight.grad.t(), **TOL)


def _check_embeddings_and_heads():
    torch.manual_seed(3)
    emb = nn.Embedding(50, 16, padding_idx=1)                   # 50 rows: padded to 64 so that 2 ranks x 32
    ids = torch.tensor([[0, 1, 31, 32, 49, 1], [7, 33, 33, 2, 48, 0]])
    vp = VocabParallelEmbedding1D.from_native_module(copy.deepcopy(emb), process_group=None,
                                                     make_vocab_size_divisible_by=32)
    assert vp.weight.shape == (32, 16)
    torch.testing.assert_close(vp(ids), emb(ids), **TOL)
    w = torch.randn(2, 6, 16)
    (vp(ids) * w).sum().backward()
    (emb(ids) * w).sum().backward()
    torch.testing.assert_close(_full_grad(vp.weight)[:50], emb.weight.grad, **TOL)
    assert float(emb.weight.grad[1].abs().sum()) == 0.0         # the padding row gets no gradient, sharded or not
    e1 = Embedding1D.from_native_module(copy.deepcopy(emb), process_group=None)      # sharded along the hidden dim
    assert e1.weight.shape == (50, 8)
    torch.testing.assert_close(e1(ids), emb(ids), **TOL)
    pe = PaddingEmbedding.from_native_module(copy.deepcopy(emb), make_vocab_size_divisible_by=64)
    assert pe.weight.shape == (64, 16)
    torch.testing.assert_close(pe(ids), emb(ids), **TOL)
    # LM heads: vocab-parallel (gathered logits are cut back to the true vocab) and padded
    head = nn.Linear(16, 50, bias=False)
    h = torch.randn(2, 6, 16)
    vh = VocabParallelLMHead1D.from_native_module(copy.deepcopy(head), process_group=None, gather_output=True,
                                                  make_vocab_size_divisible_by=32)
    torch.testing.assert_close(vh(h)[..., :50], head(h), **TOL)
    ph = PaddingLMHead.from_native_module(copy.deepcopy(head), make_vocab_size_divisible_by=64)
    torch.testing.assert_close(ph(h)[..., :50], head(h), **TOL)
    # vocab-parallel logits straight into the distributed cross entropy / log-prob (no gather of [T, V])
    local = VocabParallelLMHead1D.from_native_module(copy.deepcopy(head), process_group=None, gather_output=False,
                                                     make_vocab_size_divisible_by=32)
    labels = torch.tensor([3, 49, -100, 31, 32, 0, 7, 7, -100, 48, 1, 2])
    hr, hs = h.clone().requires_grad_(), h.clone().requires_grad_()
    loss = cross_entropy_1d(local(hs).reshape(12, -1), labels, process_group=None, vocab_size=50)
    ref = F.cross_entropy(head(hr).reshape(12, -1), labels, ignore_index=-100)
    torch.testing.assert_close(loss, ref, **TOL)
    loss.backward()
    ref.backward()
    torch.testing.assert_close(hs.grad, hr.grad, **TOL)
    torch.testing.assert_close(_full_grad(local.weight)[:50], head.weight.grad, **TOL)
    lp = dist_log_prob_1d(local(h).reshape(12, -1), labels.clamp(min=0), process_group=None, vocab_size=50)
    ref_lp = F.log_softmax(head(h).reshape(12, -1), -1).gather(-1, labels.clamp(min=0)[:, None])
    torch.testing.assert_close(lp.reshape(-1), ref_lp.reshape(-1), **TOL)


def _check_norms_and_dropout():
    torch.manual_seed(4)
    ln = nn.LayerNorm(16)
    with torch.no_grad():
        ln.weight.normal_()
        ln.bias.normal_()
    x = torch.randn(3, 5, 16)
    f = FusedLayerNorm.from_native_module(copy.deepcopy(ln), sp_partial_derived=True)
    torch.testing.assert_close(f(x), ln(x), atol=1e-5, rtol=1e-4)
    assert all(SeqParallelUtils.is_sp_partial_derived_param(p) for p in f.parameters())

    class _RMS(nn.Module):                                       # the HF-style module the policies replace
        def __init__(self):
            super().__init__()
            self.weight = nn.Parameter(torch.randn(16))
            self.variance_epsilon = 1e-6

        def forward(self, h):
            return self.weight * h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + self.variance_epsilon)

    rms = _RMS()
    fr = FusedRMSNorm.from_native_module(copy.deepcopy(rms))
    xr, xs = x.clone().requires_grad_(), x.clone().requires_grad_()
    torch.testing.assert_close(fr(xs), rms(xr), atol=1e-5, rtol=1e-4)
    fr(xs).square().sum().backward()
    rms(xr).square().sum().backward()
    torch.testing.assert_close(xs.grad, xr.grad, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(fr.weight.grad, rms.weight.grad, atol=1e-4, rtol=1e-4)
    # dropout: sharded activations draw different masks per rank, replicated ones the same mask
    torch.manual_seed(11)                                        # same seed on both ranks, as after `launch(seed=)`
    par = DropoutForParallelInput.from_native_module(nn.Dropout(0.5), process_group=None)
    rep = DropoutForReplicatedInput.from_native_module(nn.Dropout(0.5), process_group=None)
    ones = torch.ones(64, 64)
    masks = [par(ones), rep(ones)]
    both = [[torch.empty_like(ones) for _ in range(2)] for _ in masks]
    for m, b in zip(masks, both):
        dist.all_gather(b, m)
    assert not torch.equal(both[0][0], both[0][1]) and torch.equal(both[1][0], both[1][1])
    assert 0.3 < float((masks[0] == 0).float().mean()) < 0.7
    par.eval()
    assert torch.equal(par(ones), ones)


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    _check_linear_col_row()
    _check_sequence_parallel_linears()
    _check_fused_linears()
    _check_embeddings_and_heads()
    _check_norms_and_dropout()
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_parallel_layers_tp2():
    spawn(_worker, 2)
