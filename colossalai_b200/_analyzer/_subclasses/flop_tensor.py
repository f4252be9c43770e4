"""FLOP counting by operator interception: `flop_count(module, *args)` -> (forward FLOPs, backward FLOPs).

Parity: reference `colossalai/_analyzer/_subclasses/flop_tensor.py` (`flop_count`, `flop_mapping`, `Phase`).  A
`TorchDispatchMode` sees every aten call of the forward and of the autograd-driven backward; `flop_mapping` prices the
GEMM-like ones by their contraction sizes (2 * M * N * K), attention kernels by their two GEMMs, normalisations and
pointwise ops by element counts.  Runs on meta tensors, so a 70B model costs nothing to count."""
from __future__ import annotations

import math
from enum import Enum
from typing import Any, Callable, Dict, Tuple

import torch
from torch.utils._python_dispatch import TorchDispatchMode
from torch.utils._pytree import tree_map

__all__ = ["flop_count", "flop_mapping", "Phase"]

aten = torch.ops.aten


class Phase(Enum):
    FWD = 0
    BWD = 1


def _numel(shape) -> int:
    return int(math.prod(shape))


def _mm(args, out) -> int:
    a, b = args[0], args[1]
    return 2 * a.shape[0] * a.shape[1] * b.shape[1]


def _addmm(args, out) -> int:
    a, b = args[1], args[2]
    return 2 * a.shape[0] * a.shape[1] * b.shape[1] + _numel(out.shape)


def _bmm(args, out) -> int:
    a, b = args[0], args[1]
    return 2 * a.shape[0] * a.shape[1] * a.shape[2] * b.shape[2]


def _baddbmm(args, out) -> int:
    a, b = args[1], args[2]
    return 2 * a.shape[0] * a.shape[1] * a.shape[2] * b.shape[2] + _numel(out.shape)


def _conv(args, out) -> int:
    x, w = args[0], args[1]
    transposed = args[6] if len(args) > 6 else False
    o = out if isinstance(out, torch.Tensor) else out[0]
    spatial = _numel((x.shape if transposed else o.shape)[2:])
    return 2 * x.shape[0] * spatial * _numel(w.shape)


def _conv_bwd(args, out) -> int:
    grad, x, w = args[0], args[1], args[2]
    mask = args[-1]
    per = 2 * x.shape[0] * _numel(grad.shape[2:]) * _numel(w.shape)
    return per * (int(bool(mask[0])) + int(bool(mask[1])))


def _sdpa_fwd(args, out) -> int:
    q, k, v = args[0], args[1], args[2]
    B, H, Sq, D = q.shape
    return 2 * B * H * Sq * k.shape[2] * D + 2 * B * H * Sq * k.shape[2] * v.shape[3]


def _sdpa_bwd(args, out) -> int:
    q, k, v = args[1], args[2], args[3]
    B, H, Sq, D = q.shape
    # recomputed scores + dV + dP + dQ + dK: five GEMMs of 2 * B * H * Sq * Sk * D each (2.5x the forward)
    return 10 * B * H * Sq * k.shape[2] * D


def _elem(factor: float = 1.0) -> Callable:
    def f(args, out) -> int:
        o = out if isinstance(out, torch.Tensor) else next((t for t in out if isinstance(t, torch.Tensor)), None)
        return 0 if o is None else int(factor * _numel(o.shape))
    return f


def _norm(args, out) -> int:
    return 5 * _numel(args[0].shape)


def _norm_bwd(args, out) -> int:
    return 8 * _numel(args[0].shape)


flop_mapping: Dict[Any, Callable] = {
    aten.mm.default: _mm, aten.addmm.default: _addmm, aten.bmm.default: _bmm, aten.baddbmm.default: _baddbmm,
    aten.convolution.default: _conv, aten.convolution_backward.default: _conv_bwd,
    aten.native_layer_norm.default: _norm, aten.native_layer_norm_backward.default: _norm_bwd,
    aten.native_batch_norm.default: _norm, aten.native_batch_norm_backward.default: _norm_bwd,
    aten._softmax.default: _elem(5), aten._softmax_backward_data.default: _elem(4),
    aten._log_softmax.default: _elem(5), aten._log_softmax_backward_data.default: _elem(4),
}
for _name, _fn in (("_scaled_dot_product_flash_attention", _sdpa_fwd), ("_scaled_dot_product_efficient_attention", _sdpa_fwd),
                   ("_scaled_dot_product_flash_attention_for_cpu", _sdpa_fwd),
                   ("_scaled_dot_product_flash_attention_backward", _sdpa_bwd),
                   ("_scaled_dot_product_efficient_attention_backward", _sdpa_bwd),
                   ("_scaled_dot_product_flash_attention_for_cpu_backward", _sdpa_bwd)):
    if hasattr(aten, _name):
        flop_mapping[getattr(aten, _name).default] = _fn
for _name in ("add", "sub", "mul", "div", "relu", "gelu", "silu", "tanh", "sigmoid", "exp", "rsqrt", "pow", "neg",
              "threshold_backward", "gelu_backward", "silu_backward", "tanh_backward", "sigmoid_backward",
              "native_dropout", "native_dropout_backward", "hardtanh", "hardtanh_backward", "mean", "sum", "where"):
    op = getattr(aten, _name, None)
    if op is None:
        continue
    for overload in op.overloads():
        flop_mapping.setdefault(getattr(op, overload), _elem(1))


class _FlopMode(TorchDispatchMode):
    def __init__(self) -> None:
        super().__init__()
        self.phase = Phase.FWD
        self.flops = {Phase.FWD: 0, Phase.BWD: 0}
        self.by_op: Dict[str, int] = {}

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        fn = flop_mapping.get(func)
        if fn is not None:
            try:
                n = int(fn(args, out))
            except Exception:                               # an overload with an unusual signature: do not price it
                n = 0
            self.flops[self.phase] += n
            self.by_op[str(func)] = self.by_op.get(str(func), 0) + n
        return out


def flop_count(module, *args, verbose: bool = False, **kwargs) -> Tuple[int, int]:
    """FLOPs of one forward and one backward of `module(*args, **kwargs)` (a module or any callable).  Tensors are
    moved to the meta device first; a module is deep-copied to meta so the original stays untouched."""
    import copy

    def to_meta(t):
        if isinstance(t, torch.Tensor):
            base = getattr(t, "_tensor", t)
            m = base.detach().to("meta")
            return m.requires_grad_(True) if m.is_floating_point() else m
        return t

    fn = module
    if isinstance(module, torch.nn.Module):
        fn = copy.deepcopy(module).to("meta")
        fn.train(module.training)
    args, kwargs = tree_map(to_meta, args), tree_map(to_meta, kwargs)
    mode = _FlopMode()
    with mode:
        out = fn(*args, **kwargs)
        mode.phase = Phase.BWD
        outs = []
        tree_map(lambda t: outs.append(t) if isinstance(t, torch.Tensor) and t.requires_grad else None, out)
        if outs:
            torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])
    if verbose:
        width = max((len(k) for k in mode.by_op), default=10)
        for k, v in sorted(mode.by_op.items(), key=lambda kv: -kv[1]):
            print(f"{k:{width}s} {v / 1e9:12.3f} GFLOP")
    return mode.flops[Phase.FWD], mode.flops[Phase.BWD]
