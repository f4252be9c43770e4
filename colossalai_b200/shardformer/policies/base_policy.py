"""Policy ABC + replacement descriptions.  Parity: reference `colossalai/shardformer/policies/base_policy.py:19-209`."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Type, Union

import torch.nn as nn
from torch import Tensor

from ..shard.shard_config import ShardConfig

__all__ = ["ParallelModule", "SubModuleReplacementDescription", "ModulePolicyDescription", "Policy"]


@dataclass
class SubModuleReplacementDescription:
    """Replace sub-module `suffix` (dotted path relative to the matched module) by
    `target_module.from_native_module(original, process_group, **kwargs)`."""

    suffix: str
    target_module: Union[Type[nn.Module], Callable]
    kwargs: Dict[str, Any] = None
    ignore_if_not_exist: bool = False


@dataclass
class ModulePolicyDescription:
    attribute_replacement: Dict[str, Any] = None
    param_replacement: List[Callable] = None
    sub_module_replacement: List[SubModuleReplacementDescription] = None
    method_replacement: Dict[str, Callable] = None


class Policy(ABC):
    """Describes how to shard one model family."""

    def __init__(self) -> None:
        self.shard_config: Optional[ShardConfig] = None
        self.model: Optional[nn.Module] = None
        self.is_causal = None

    def set_model(self, model: nn.Module) -> None:
        self.model = model

    def set_shard_config(self, shard_config: ShardConfig) -> None:
        self.shard_config = shard_config
        self.config_sanity_check()

    @property
    def pipeline_stage_manager(self):
        return self.shard_config.pipeline_stage_manager if self.shard_config is not None else None

    @abstractmethod
    def config_sanity_check(self) -> None:
        ...

    @abstractmethod
    def preprocess(self) -> nn.Module:
        ...

    @abstractmethod
    def module_policy(self) -> Dict[Union[str, Type[nn.Module]], ModulePolicyDescription]:
        ...

    @abstractmethod
    def postprocess(self) -> nn.Module:
        ...

    def get_held_layers(self) -> List[nn.Module]:
        """Modules kept by this pipeline stage (everything when PP is off)."""
        return [self.model]

    def get_shared_params(self) -> List[Dict[int, Tensor]]:
        """[{stage: param, ...}] for parameters tied across pipeline stages."""
        return []

    def append_or_create_submodule_replacement(
        self, description: Union[SubModuleReplacementDescription, List[SubModuleReplacementDescription]],
        policy: Dict[Union[str, Type[nn.Module]], ModulePolicyDescription], target_key: Union[str, Type[nn.Module]],
    ) -> Dict:
        if isinstance(description, SubModuleReplacementDescription):
            description = [description]
        if target_key in policy:
            if policy[target_key].sub_module_replacement is None:
                policy[target_key].sub_module_replacement = list(description)
            else:
                policy[target_key].sub_module_replacement.extend(description)
        else:
            policy[target_key] = ModulePolicyDescription(sub_module_replacement=list(description))
        return policy

    def append_or_create_method_replacement(self, description: Dict[str, Callable], policy: Dict,
                                            target_key: Union[str, Type[nn.Module]]) -> Dict:
        if target_key in policy:
            if policy[target_key].method_replacement is None:
                policy[target_key].method_replacement = dict(description)
            else:
                policy[target_key].method_replacement.update(description)
        else:
            policy[target_key] = ModulePolicyDescription(method_replacement=dict(description))
        return policy

    def tie_weight_check(self) -> bool:
        in_emb = getattr(self.model, "get_input_embeddings", lambda: None)()
        out_emb = getattr(self.model, "get_output_embeddings", lambda: None)()
        return (in_emb is not None and out_emb is not None and id(in_emb.weight) == id(out_emb.weight))
