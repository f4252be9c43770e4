"""Diffusion pipelines behind the same engine façade.
Parity: reference `colossalai/inference/core/diffusion_engine.py:27-200` (PixArt-alpha / Stable-Diffusion-3 through
`diffusers` pipelines with a request queue).  `diffusers` is an optional dependency: the engine accepts any object with
the diffusers pipeline protocol (`transformer`/`unet`, `vae`, `__call__(prompt=..., **kw) -> .images`)."""
from __future__ import annotations

from itertools import count
from typing import List, Optional, Union

import torch

from ..config import InferenceConfig
from .base_engine import BaseEngine

__all__ = ["DiffusionEngine"]


class DiffusionEngine(BaseEngine):
    def __init__(self, model_or_path, inference_config: Optional[InferenceConfig] = None, verbose: bool = False,
                 model_policy=None) -> None:
        self.inference_config = inference_config or InferenceConfig()
        self.dtype = self.inference_config.dtype
        self.verbose = verbose
        self.counter = count()
        self._queue: List[dict] = []
        self.init_model(model_or_path, model_policy)
        self._verify_args()

    def init_model(self, model_or_path, model_policy=None, model_shard_infer_config=None) -> None:
        if isinstance(model_or_path, str):
            try:
                from diffusers import DiffusionPipeline
            except ImportError as e:  # pragma: no cover - optional dependency
                raise ImportError("DiffusionEngine needs the optional `diffusers` package to load a pipeline from a "
                                  "path; pass an already constructed pipeline object instead") from e
            self.model = DiffusionPipeline.from_pretrained(model_or_path, torch_dtype=self.dtype)
        else:
            self.model = model_or_path
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        if hasattr(self.model, "to"):
            self.model = self.model.to(dev)

    def _verify_args(self) -> None:
        assert callable(self.model), "a diffusion pipeline must be callable"

    def add_request(self, prompts: Union[List[str], str], request_ids=None, **kwargs) -> None:
        if isinstance(prompts, str):
            prompts = [prompts]
        for i, p in enumerate(prompts):
            rid = request_ids[i] if request_ids else next(self.counter)
            self._queue.append({"request_id": rid, "prompt": p, "kwargs": kwargs})

    def step(self):
        if not self._queue:
            return []
        req = self._queue.pop(0)
        with torch.inference_mode():
            out = self.model(prompt=req["prompt"], **req["kwargs"])
        return [(req["request_id"], getattr(out, "images", out))]

    def generate(self, request_ids=None, prompts=None, generation_config=None, **kwargs):
        if prompts is not None:
            self.add_request(prompts, request_ids, **kwargs)
        results = []
        while self._queue:
            results += self.step()
        return [img for _, img in sorted(results, key=lambda r: r[0])]
