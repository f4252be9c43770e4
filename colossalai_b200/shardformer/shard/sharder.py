"""ModelSharder: apply a policy to a model.

Parity: reference `colossalai/shardformer/shard/sharder.py:18-236`: preprocess -> shared params -> release layers
not held by this PP stage -> recursive replace (attributes, params, methods, sub-modules) -> materialise lazy
tensors -> postprocess.
"""
from __future__ import annotations

from types import MethodType
from typing import Any, Callable, Dict, List, Optional, Set, Tuple, Union

import torch.nn as nn
from torch import Tensor

from .._utils import getattr_, set_tensors_to_none, setattr_
from ..policies.auto_policy import get_autopolicy
from ..policies.base_policy import Policy, SubModuleReplacementDescription
from .shard_config import ShardConfig

__all__ = ["ModelSharder", "shard_model"]


class ModelSharder:
    def __init__(self, model: nn.Module, policy: Optional[Policy], shard_config: ShardConfig = None) -> None:
        self.model = model
        self.shard_config = shard_config
        self.policy = get_autopolicy(self.model) if policy is None else policy

    def shard(self) -> List[Dict[int, Tensor]]:
        self.policy.set_model(self.model)
        self.policy.set_shard_config(self.shard_config)
        self._preprocess()
        shared_params = self.policy.get_shared_params()   # before releasing layers
        held_layers = self._release_unheld_layers()
        self._replace_module(include=held_layers)
        self._materialize()
        self._postprocess()
        return shared_params

    def _preprocess(self) -> None:
        self.model = self.policy.preprocess()

    def _postprocess(self) -> None:
        self.model = self.policy.postprocess()

    def _replace_module(self, include: Optional[Set[nn.Module]] = None) -> None:
        for layer_cls, desc in self.policy.module_policy().items():
            self._recursive_replace_layer(self.model, layer_cls, desc.attribute_replacement, desc.param_replacement,
                                          desc.method_replacement, desc.sub_module_replacement, include=include)

    def _recursive_replace_layer(self, module: nn.Module, origin_cls: Union[str, type],
                                 attr_replacement: Dict[str, Any], param_replacement: List[Callable],
                                 method_replacement: Dict[str, Callable],
                                 sub_module_replacement: List[SubModuleReplacementDescription],
                                 include: Optional[Set[nn.Module]] = None) -> None:
        matched = (isinstance(origin_cls, str) and origin_cls == module.__class__.__name__) or \
                  (isinstance(origin_cls, type) and isinstance(module, origin_cls))
        if matched:
            if attr_replacement is not None:
                for k, v in attr_replacement.items():
                    setattr_(module, k, v)
            if param_replacement is not None and (include is None or module in include):
                for fn in param_replacement:
                    fn(module)
            if method_replacement is not None:
                for name, fn in method_replacement.items():
                    setattr(module, name, MethodType(fn, module))
            if sub_module_replacement is not None:
                self._replace_sub_module(module, sub_module_replacement, include)
        for child in module.children():
            self._recursive_replace_layer(child, origin_cls, attr_replacement, param_replacement, method_replacement,
                                          sub_module_replacement, include=include)

    def _replace_sub_module(self, org_layer: nn.Module, sub_module_replacement: List[SubModuleReplacementDescription],
                            include: Optional[Set[nn.Module]] = None) -> None:
        for desc in sub_module_replacement:
            suffix, target, kwargs = desc.suffix, desc.target_module, desc.kwargs or {}
            native = getattr_(org_layer, suffix, ignore=True)
            if native is None:
                if desc.ignore_if_not_exist:
                    continue
                raise AttributeError(f"{org_layer.__class__.__name__} has no sub-module {suffix!r}")
            if getattr(native, "_cb200_replaced_by", None) is target:   # shared module already converted
                continue
            if include is not None and native not in include:
                continue
            try:
                replaced = target.from_native_module(native, process_group=self.shard_config.tensor_parallel_process_group,
                                                     **kwargs)
            except Exception as e:
                raise RuntimeError(
                    f"failed to replace {suffix} of type {native.__class__.__qualname__} with "
                    f"{getattr(target, '__qualname__', target)}: {e}") from e
            try:
                replaced._cb200_replaced_by = target
            except Exception:
                pass
            setattr_(org_layer, suffix, replaced)

    def _get_recursive_held_layers(self, held_layers: Optional[List[nn.Module]]) -> Optional[List[nn.Module]]:
        def collect(module: nn.Module, acc: List[nn.Module]):
            acc.append(module)
            for c in module.children():
                collect(c, acc)

        if held_layers is None:
            return None
        out: List[nn.Module] = []
        for m in held_layers:
            collect(m, out)
        return out

    def _release_unheld_layers(self) -> Optional[Set[nn.Module]]:
        if self.shard_config is not None and self.shard_config.pipeline_stage_manager is not None:
            held = self.policy.get_held_layers()
            set_tensors_to_none(self.model, exclude=set(held))
            return set(self._get_recursive_held_layers(held))
        return None

    def _materialize(self) -> None:
        from ...lazy import LazyInitContext

        LazyInitContext.materialize(self.model)


def shard_model(model: nn.Module, shard_config: ShardConfig = None, policy: Policy = None):
    sharder = ModelSharder(model=model, shard_config=shard_config, policy=policy)
    shared = sharder.shard()
    return model, shared
