"""InferenceEngine façade.  Parity: reference `colossalai/inference/core/engine.py:16-133`."""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from ..config import GenerationConfig, InferenceConfig
from .llm_engine import LLMEngine

__all__ = ["InferenceEngine"]


class InferenceEngine:
    """`InferenceEngine(model_or_path, tokenizer, inference_config).generate(prompts=[...])`"""

    def __init__(self, model_or_path, tokenizer=None, inference_config: InferenceConfig = None, verbose: bool = False,
                 model_policy=None) -> None:
        self.__dict__["_initialized"] = False
        self.engine = None
        self.tokenizer = tokenizer
        self.inference_config = inference_config or InferenceConfig()
        pipeline_like = hasattr(model_or_path, "transformer") and hasattr(model_or_path, "vae")
        if pipeline_like:
            from .diffusion_engine import DiffusionEngine

            self.engine = DiffusionEngine(model_or_path, inference_config=self.inference_config, verbose=verbose,
                                          model_policy=model_policy)
        else:
            self.engine = LLMEngine(model_or_path, tokenizer, self.inference_config, verbose, model_policy)
        self.__dict__["_initialized"] = True

    def _verify_args(self) -> None:
        assert self.engine is not None

    def generate(self, request_ids: Union[List[int], int] = None, prompts: Union[List[str], str] = None, *args,
                 **kwargs):
        assert self.engine is not None, "Please init Engine first"
        return self.engine.generate(request_ids=request_ids, prompts=prompts, *args, **kwargs)

    def add_request(self, request_ids: Union[List[int], int] = None, prompts: Union[List[str], str] = None, *args,
                    **kwargs) -> None:
        self.engine.add_request(request_ids=request_ids, prompts=prompts, *args, **kwargs)

    def step(self):
        return self.engine.step()

    def __getattr__(self, name):
        if self.__dict__.get("_initialized"):
            return getattr(self.__dict__["engine"], name)
        raise AttributeError(name)
