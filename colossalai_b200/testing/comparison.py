"""Numeric comparison helpers.  Parity: reference `colossalai/testing/comparison.py`."""
from __future__ import annotations

from typing import Any, List, Optional, OrderedDict

import torch
import torch.distributed as dist
from torch import Tensor
from torch.distributed import ProcessGroup
from torch.testing import assert_close as _torch_assert_close


def assert_equal(a: Tensor, b: Tensor) -> None:
    assert torch.all(a == b), f"expected a and b to be equal but they are not, {a} vs {b}"


def assert_not_equal(a: Tensor, b: Tensor) -> None:
    assert not torch.all(a == b), "expected a and b to differ"


def assert_close(a: Tensor, b: Tensor, rtol: float = 1e-5, atol: float = 1e-8, **kw) -> None:
    _torch_assert_close(a, b.to(a.dtype).to(a.device), rtol=rtol, atol=atol, **kw)


def assert_close_loose(a: Tensor, b: Tensor, rtol: float = 1e-3, atol: float = 1e-3) -> None:
    assert_close(a, b, rtol=rtol, atol=atol)


def assert_equal_in_group(tensor: Tensor, process_group: Optional[ProcessGroup] = None) -> None:
    world = dist.get_world_size(process_group)
    gathered = [torch.empty_like(tensor) for _ in range(world)]
    dist.all_gather(gathered, tensor, group=process_group)
    for i in range(world - 1):
        assert torch.all(gathered[i] == gathered[i + 1]), f"rank {i} and {i + 1} differ"


def check_state_dict_equal(d1: OrderedDict, d2: OrderedDict, ignore_device: bool = True,
                           ignore_dtype: bool = False) -> None:
    assert len(list(d1.keys())) == len(list(d2.keys())), f"#keys differ: {len(d1)} vs {len(d2)}"
    for k, v1 in d1.items():
        assert k in d2, f"key {k} missing"
        v2 = d2[k]
        _check(k, v1, v2, ignore_device, ignore_dtype)


def _check(k: Any, v1: Any, v2: Any, ignore_device: bool, ignore_dtype: bool) -> None:
    if isinstance(v1, dict):
        assert isinstance(v2, dict)
        for kk in v1:
            _check(f"{k}.{kk}", v1[kk], v2[kk], ignore_device, ignore_dtype)
    elif isinstance(v1, (list, tuple)):
        assert len(v1) == len(v2)
        for i, (a, b) in enumerate(zip(v1, v2)):
            _check(f"{k}[{i}]", a, b, ignore_device, ignore_dtype)
    elif isinstance(v1, Tensor):
        assert isinstance(v2, Tensor), f"{k}: tensor vs {type(v2)}"
        if not ignore_device:
            v2 = v2.to(v1.device)
        else:
            v1, v2 = v1.cpu(), v2.cpu()
        if ignore_dtype:
            v2 = v2.to(v1.dtype)
        assert v1.shape == v2.shape, f"{k}: shape {v1.shape} vs {v2.shape}"
        assert v1.dtype == v2.dtype, f"{k}: dtype {v1.dtype} vs {v2.dtype}"
        _torch_assert_close(v1, v2, rtol=3e-3, atol=3e-3, msg=lambda m: f"{k}: {m}")
    else:
        assert v1 == v2, f"{k}: {v1} vs {v2}"


def assert_hf_output_close(out1: Any, out2: Any, ignore_keys: List[str] = None, track_name: str = "",
                           atol: float = 1e-5, rtol: float = 1e-5) -> None:
    if isinstance(out1, dict) and isinstance(out2, dict):
        for k in out1.keys():
            if ignore_keys is not None and k in ignore_keys:
                continue
            assert_hf_output_close(out1[k], out2[k], ignore_keys, f"{track_name}.{k}", atol, rtol)
    elif isinstance(out1, (list, tuple)) and isinstance(out2, (list, tuple)):
        for i in range(len(out1)):
            assert_hf_output_close(out1[i], out2[i], ignore_keys, f"{track_name}.{i}", atol, rtol)
    elif isinstance(out1, Tensor) and isinstance(out2, Tensor):
        assert out1.shape == out2.shape, f"{track_name}: {out1.shape} vs {out2.shape}"
        _torch_assert_close(out1, out2.to(out1.dtype), atol=atol, rtol=rtol, msg=lambda m: f"{track_name}: {m}")
    else:
        assert out1 == out2, f"{track_name}: {out1} vs {out2}"
