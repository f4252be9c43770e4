"""Lazy (meta-device) model construction: build -> shard -> materialise only the local shards.

Parity: reference `colossalai/lazy/lazy_init.py:134-603` (`LazyTensor` op log + `LazyInitContext.materialize`)
and `lazy/pretrained.py` (deferred `from_pretrained`).  Design here: parameters are created on the `meta` device
and a TorchFunctionMode records every in-place initialiser call (`normal_`, `uniform_`, `zero_`, `fill_`, ...)
on them; parallel layers built from a meta module inherit the op log (initialisers are shape-agnostic), and
`materialize` allocates the *local shard* on the target device and replays the log under the layer's per-rank RNG.
This lets a 70B model be constructed on 8 GPUs without ever holding a full weight anywhere.
`materialize(module, reproduce_eager=True)` is the reproducibility path: a second, complete log (constructor defaults
through `torch.nn.init.*`, tensors the model dropped again) is replayed for the whole model in recording order, which
consumes the RNG exactly like eager construction - lazy and eager builds are then equal value for value
(reference `tests/test_lazy/test_models.py`).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.nn as nn
from torch.overrides import TorchFunctionMode

__all__ = ["LazyInitContext", "LazyTensor", "copy_lazy_ops", "is_lazy"]

_INPLACE_INIT = {"normal_", "uniform_", "zero_", "fill_", "fill_diagonal_", "copy_", "trunc_normal_", "mul_", "add_",
                 "div_", "clamp_", "bernoulli_", "random_", "exponential_", "erfinv_", "sub_"}


class LazyTensor:
    """Marker namespace kept for API parity (lazy tensors are plain meta tensors carrying `_lazy_ops`)."""

    @staticmethod
    def is_lazy(t: torch.Tensor) -> bool:
        return is_lazy(t)


def is_lazy(t: torch.Tensor) -> bool:
    return isinstance(t, torch.Tensor) and t.device.type == "meta"


def copy_lazy_ops(src: Optional[torch.Tensor], dst: Optional[torch.Tensor]) -> None:
    if src is not None and dst is not None and hasattr(src, "_lazy_ops"):
        dst._lazy_ops = list(src._lazy_ops)


class _Recorder(TorchFunctionMode):
    """Two logs per meta tensor.  `_lazy_ops`: the tensor-level in-place initialisers, replayed shard by shard.
    `_lazy_full_log`: EVERY random-or-not initialiser that touched the tensor, `torch.nn.init.*` defaults of the module
    constructors included, each with a global sequence number - replaying all tensors' full logs in that order consumes
    the RNG exactly like eager construction did (`materialize(..., reproduce_eager=True)`)."""

    def __init__(self) -> None:
        super().__init__()
        self.seq = 0
        self.all_tensors: List[torch.Tensor] = []      # every logged meta tensor, also ones the model later drops

    def _log_full(self, t: torch.Tensor, name: str, rest, kwargs) -> None:
        log = getattr(t, "_lazy_full_log", None)
        if log is None:
            log = []
            try:
                t._lazy_full_log = log
                t._lazy_all_tensors = self.all_tensors
            except Exception:
                return
            self.all_tensors.append(t)
        log.append((self.seq, name, rest, dict(kwargs)))
        self.seq += 1

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        out = func(*args, **kwargs)
        name = getattr(func, "__name__", "")
        if getattr(func, "__module__", "") == "torch.nn.init" and hasattr(nn.init, name):
            # the wrappers dispatch with the tensor as keyword (`tensor=`) or first positional argument
            t = kwargs.get("tensor", args[0] if args else None)
            if isinstance(t, torch.Tensor) and t.device.type == "meta":
                kw = {k: v for k, v in kwargs.items() if k != "tensor"}
                self._log_full(t, "init:" + name, tuple(args[1:]) if args else (), kw)
        elif name in _INPLACE_INIT and args and isinstance(args[0], torch.Tensor) and args[0].device.type == "meta" \
                and not any(isinstance(a, torch.Tensor) and a.device.type == "meta" for a in args[1:]):
            self._log_full(args[0], name, tuple(args[1:]), kwargs)
        # NB: `torch.nn.init.*` wrappers dispatch as a whole (name e.g. "kaiming_uniform_") and are deliberately NOT
        # recorded: module constructors call them as throw-away defaults.  Only tensor-level in-place initialisers
        # (`w.normal_()`, `w.fill_()`, ...) are replayed; everything else gets `_default_fill` (or the model's own
        # `_init_weights` after materialisation).
        if getattr(func, "__module__", "") != "torch.nn.init" and name in _INPLACE_INIT and args \
                and isinstance(args[0], torch.Tensor) and args[0].device.type == "meta":
            t = args[0]
            rest = tuple(a for a in args[1:])
            if any(isinstance(a, torch.Tensor) and a.device.type == "meta" for a in rest):
                return out  # cannot replay ops that depend on other lazy tensors
            ops = getattr(t, "_lazy_ops", None)
            if ops is None:
                ops = []
                try:
                    t._lazy_ops = ops
                except Exception:
                    return out
            ops.append((name, rest, dict(kwargs)))
        return out


class LazyInitContext:
    """`with LazyInitContext(): model = Model(cfg)` then `ShardFormer.optimize` / `booster.boost` materialises."""

    _replaced = False

    def __init__(self, tensor_cls=None, default_device: Optional[torch.device] = None) -> None:
        self.default_device = default_device
        self._dev_ctx = None
        self._rec = None

    def __enter__(self):
        self._dev_ctx = torch.device("meta")
        self._dev_ctx.__enter__()
        self._rec = _Recorder()
        self._rec.__enter__()
        return self

    def __exit__(self, exc_type, exc, tb):
        self._rec.__exit__(exc_type, exc, tb)
        self._dev_ctx.__exit__(exc_type, exc, tb)

    # ------------------------------------------------------------------ materialise
    @staticmethod
    def _default_fill(name: str, t: torch.Tensor) -> None:
        if t.dim() >= 2:
            nn.init.normal_(t, std=0.02)
        elif "norm" in name and name.endswith("weight"):
            nn.init.ones_(t)
        else:
            nn.init.zeros_(t)

    @staticmethod
    def _replay(name: str, meta_t: torch.Tensor, real: torch.Tensor) -> None:
        ops = getattr(meta_t, "_lazy_ops", None)
        if not ops:
            LazyInitContext._default_fill(name, real)
            return
        with torch.no_grad():
            for fn, args, kwargs in ops:
                if fn.startswith("init:"):
                    getattr(nn.init, fn[5:])(real, *args, **kwargs)
                elif fn == "trunc_normal_":
                    nn.init.trunc_normal_(real, *args, **kwargs)
                else:
                    getattr(real, fn)(*args, **kwargs)

    @staticmethod
    def _materialize_like_eager(module: nn.Module, device) -> None:
        """Allocate ALL meta parameters / buffers and replay their full logs in the global order they were recorded in:
        under the seed the context was entered with this yields bit-identical values to eager construction (needs the
        whole model on `device` at once - it is the debugging / reproducibility path, not the 70B one)."""
        entries, seen = [], {}
        for mod_name, mod in module.named_modules():
            for store in (mod._parameters, mod._buffers):
                for name, t in list(store.items()):
                    if t is None or t.device.type != "meta":
                        continue
                    if id(t) not in seen:
                        real = torch.zeros(t.shape, dtype=t.dtype, device=device)
                        if isinstance(t, nn.Parameter):
                            newp = nn.Parameter(real, requires_grad=t.requires_grad)
                            for attr, val in vars(t).items():
                                if attr not in ("_lazy_ops", "_lazy_full_log"):      # (`_lazy_all_tensors` is copied on purpose)
                                    try:
                                        setattr(newp, attr, val)
                                    except Exception:
                                        pass
                            real = newp
                        seen[id(t)] = real
                        full = f"{mod_name}.{name}" if mod_name else name
                        log = getattr(t, "_lazy_full_log", None)
                        if log:
                            entries += [(seq, real, fn, a, k) for seq, fn, a, k in log]
                        elif isinstance(t, nn.Parameter):
                            entries.append((float("inf"), real, "default:" + full, (), {}))
                    store[name] = seen[id(t)]
        # tensors that were initialised during construction but are no longer part of the model (a head's own weight
        # replaced by the tied embedding, ...) consumed random numbers in the eager run: burn the same ones on scratch
        registry = None
        for real in seen.values():
            registry = vars(real).pop("_lazy_all_tensors", registry) if isinstance(real, nn.Parameter) else registry
        if registry is not None:
            for t in registry:
                if id(t) not in seen:
                    scratch = torch.zeros(t.shape, dtype=t.dtype, device=device)
                    entries += [(seq, scratch, fn, a, k) for seq, fn, a, k in t._lazy_full_log]
        with torch.no_grad():
            for _, real, fn, a, k in sorted(entries, key=lambda e: e[0]):
                data = real.data
                if fn.startswith("init:"):
                    getattr(nn.init, fn[5:])(data, *a, **k)
                elif fn.startswith("default:"):
                    LazyInitContext._default_fill(fn[8:], data)
                else:
                    getattr(data, fn)(*a, **k)

    @staticmethod
    def materialize(module: nn.Module, device: Optional[torch.device] = None, verbose: bool = False,
                    reproduce_eager: bool = False) -> nn.Module:
        """Allocate every meta parameter/buffer of `module` on `device` and replay its initialiser log.
        `reproduce_eager`: replay the full log of the whole model in recording order (same RNG stream as eager
        construction under the same seed) instead of per-module, shard-friendly replay."""
        if device is None:
            from ..accelerator import get_accelerator

            device = get_accelerator().get_current_device()
        if reproduce_eager:
            LazyInitContext._materialize_like_eager(module, device)
            return module
        memo = {}
        n = 0
        for mod_name, mod in module.named_modules():
            rnd = getattr(mod, "randomizer", None)
            for pname, p in list(mod._parameters.items()):
                if p is None or p.device.type != "meta":
                    continue
                if id(p) in memo:
                    mod._parameters[pname] = memo[id(p)]
                    continue
                real = torch.empty(p.shape, dtype=p.dtype, device=device)
                full = f"{mod_name}.{pname}" if mod_name else pname
                if rnd is not None:
                    with rnd.fork_rng(enable_cpu=torch.device(device).type == "cpu"):
                        LazyInitContext._replay(full, p, real)
                else:
                    LazyInitContext._replay(full, p, real)
                newp = nn.Parameter(real, requires_grad=p.requires_grad)
                for attr, val in vars(p).items():
                    if attr in ("_lazy_ops", "_lazy_full_log", "_lazy_all_tensors"):
                        continue
                    try:
                        setattr(newp, attr, val)
                    except Exception:
                        pass
                # re-install metadata-preserving detach/clone on the real tensor
                for a in ("_old_detach", "_old_clone", "detach", "clone", "_pad_old_detach", "_pad_old_clone"):
                    if a in vars(newp):
                        delattr(newp, a)
                if hasattr(newp, "dist_shard") or hasattr(newp, "shard_fn") or hasattr(newp, "dist_layout"):
                    from ..tensor.d_tensor.api import _hijack_detach_and_clone

                    _hijack_detach_and_clone(newp)
                if hasattr(newp, "_padding_dim"):
                    from ..tensor.padded_tensor import _hijack

                    _hijack(newp)
                zp = getattr(mod, "_zero_padding", None)
                if pname == "weight" and callable(zp):
                    zp(newp.data)
                memo[id(p)] = newp
                mod._parameters[pname] = newp
                n += 1
            for bname, b in list(mod._buffers.items()):
                if b is None or b.device.type != "meta":
                    continue
                real = torch.zeros(b.shape, dtype=b.dtype, device=device)
                LazyInitContext._replay(bname, b, real) if getattr(b, "_lazy_ops", None) else None
                mod._buffers[bname] = real
        if verbose:
            print(f"LazyInitContext.materialize: allocated {n} parameters on {device}")
        return module
