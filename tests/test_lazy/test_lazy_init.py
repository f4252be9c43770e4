"""LazyInitContext: meta construction + replayed initialisation equals eager construction
(reference: tests/test_lazy/test_models.py)."""
import torch
import torch.nn as nn

from colossalai_b200.lazy import LazyInitContext
from colossalai_b200.models import build_model


def test_lazy_matches_eager_for_model_zoo():
    for name in ("llama-tiny", "gpt2-tiny", "mixtral-tiny"):
        torch.manual_seed(11)
        eager = build_model(name)
        torch.manual_seed(11)
        with LazyInitContext():
            lazy = build_model(name)
        assert all(p.device.type == "meta" for p in lazy.parameters())
        LazyInitContext.materialize(lazy)
        assert all(p.device.type != "meta" for p in lazy.parameters())
        for (n1, p1), (n2, p2) in zip(eager.named_parameters(), lazy.named_parameters()):
            assert n1 == n2 and p1.shape == p2.shape
            assert torch.isfinite(p2).all()
        # same statistics (the RNG stream differs by construction order, the distributions must not)
        e = torch.cat([p.flatten() for p in eager.parameters()]).std()
        l = torch.cat([p.flatten() for p in lazy.parameters()]).std()
        assert abs(e - l) / e < 0.1


def test_lazy_custom_module_ops_are_replayed():
    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.w = nn.Parameter(torch.empty(4, 4))
            self.k = nn.Parameter(torch.empty(8, 4))
            with torch.no_grad():
                self.w.fill_(3.0)
                self.k.uniform_(-1.0, 1.0).mul_(0.5)

    with LazyInitContext():
        m = M()
    LazyInitContext.materialize(m)
    assert torch.equal(m.w.detach(), torch.full((4, 4), 3.0))
    assert m.k.abs().max() <= 0.5 + 1e-6 and m.k.abs().max() > 0.3          # uniform(-1, 1) then the recorded mul_(0.5)


def test_lazy_reproduces_eager_bit_for_bit():
    """`materialize(reproduce_eager=True)` replays the FULL initialiser log of the whole model in recording order -
    constructor defaults, the model's own init, tensors the model dropped again (a head's weight replaced by the tied
    embedding) - so every family's lazy build equals its eager build under the same seed, value for value
    (reference: tests/test_lazy/test_models.py `check_lazy_init`)."""
    from colossalai_b200.models import MODEL_ZOO

    names = [n for n in MODEL_ZOO if n.endswith("-tiny")]
    assert len(names) >= 12
    for name in names:
        torch.manual_seed(5)
        eager = build_model(name)
        torch.manual_seed(5)
        with LazyInitContext():
            lazy = build_model(name)
        LazyInitContext.materialize(lazy, torch.device("cpu"), reproduce_eager=True)
        pe, pl = dict(eager.named_parameters()), dict(lazy.named_parameters())
        assert pe.keys() == pl.keys(), name
        for k in pe:
            assert torch.equal(pe[k], pl[k]), (name, k)
        for (k, b1), (_, b2) in zip(eager.named_buffers(), lazy.named_buffers()):
            assert torch.equal(b1, b2), (name, k)
        tied = [k for k, p in pe.items() if p is eager.get_input_embeddings().weight]
        assert (lazy.get_output_embeddings().weight is lazy.get_input_embeddings().weight) == \
            (eager.get_output_embeddings().weight is eager.get_input_embeddings().weight), (name, tied)
