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
        from ...models.dit import DIT_ZOO, DiTConfig, build_diffusion_pipeline

        if isinstance(model_or_path, DiTConfig) or (isinstance(model_or_path, str) and model_or_path in DIT_ZOO):
            # native DiT backbones (PixArt-alpha / SD3 shapes) — no diffusers needed
            self.model = build_diffusion_pipeline(model_or_path)
            self.model.transformer.to(self.dtype)
        elif isinstance(model_or_path, str):
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
        # Distrifusion patch parallelism over the ranks of the default group
        pps = getattr(self.inference_config, "patched_parallelism_size", 1)
        self.patch_ctx = None
        if pps > 1:
            import torch.distributed as dist

            from ..modeling.layers import enable_patch_parallel

            assert dist.is_initialized() and dist.get_world_size() % pps == 0
            rank = dist.get_rank()
            group = None
            for start in range(0, dist.get_world_size(), pps):
                g = dist.new_group(list(range(start, start + pps)))
                if start <= rank < start + pps:
                    group = g
            transformer = getattr(self.model, "transformer", None)
            assert transformer is not None and hasattr(transformer, "patch_parallel"), \
                "patch parallelism needs a native DiT pipeline (models.dit)"
            self.patch_ctx = enable_patch_parallel(transformer, group,
                                                   warmup_steps=getattr(self.inference_config, "pp_warmup_steps", 4),
                                                   mode=getattr(self.inference_config, "pp_mode", "stale"))

    def _verify_args(self) -> None:
        assert callable(self.model), "a diffusion pipeline must be callable"

    def add_request(self, prompts: Union[List[str], str], request_ids=None, **kwargs) -> None:
        if isinstance(prompts, str) or (torch.is_tensor(prompts) and prompts.dim() == 3):
            prompts = [prompts]
        for i, p in enumerate(prompts):
            rid = request_ids[i] if request_ids else next(self.counter)
            self._queue.append({"request_id": rid, "prompt": p, "kwargs": kwargs})

    def step(self):
        if not self._queue:
            return []
        req = self._queue.pop(0)
        with torch.inference_mode():
            kw = dict(req["kwargs"])
            if torch.is_tensor(req["prompt"]):       # pre-computed text embeddings
                kw["prompt_embeds"] = req["prompt"]
                out = self.model(**kw)
            else:
                out = self.model(prompt=req["prompt"], **kw)
        return [(req["request_id"], getattr(out, "images", out))]

    def generate(self, request_ids=None, prompts=None, generation_config=None, **kwargs):
        if prompts is not None:
            self.add_request(prompts, request_ids, **kwargs)
        results = []
        while self._queue:
            results += self.step()
        return [img for _, img in sorted(results, key=lambda r: r[0])]
