"""Rollout producers.  A producer owns an inference copy of the policy, samples `num_generations` responses per
prompt and hands token-id rollouts to the consumer; its weights are refreshed from the consumer every
`sync_every` updates.

Two back ends: `ModelRolloutBackend` (plain sampling loop on any module) and `EngineRolloutBackend` (our paged-KV
continuous-batching `InferenceEngine`, the counterpart of the reference's vLLM backend).
Parity: reference `coati/distributed/{producer.py:1-500, inference_backend.py:1-300}`."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from ..models import generate


class ModelRolloutBackend:
    def __init__(self, model: nn.Module, generate_kwargs: Optional[Dict] = None, pad_token_id: int = 0,
                 eos_token_id: Optional[int] = None) -> None:
        self.model = model.eval()
        self.kw = dict(generate_kwargs or {})
        self.pad_token_id, self.eos_token_id = pad_token_id, eos_token_id

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, num_generations: int) -> torch.Tensor:
        ids = input_ids.repeat_interleave(num_generations, 0)
        am = attention_mask.repeat_interleave(num_generations, 0)
        return generate(self.model, ids, am, pad_token_id=self.pad_token_id, eos_token_id=self.eos_token_id, **self.kw)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.model.load_state_dict(sd, strict=False)


class EngineRolloutBackend:
    """Sampling through `colossalai_b200.inference.InferenceEngine` (paged KV, continuous batching, CUDA graphs)."""

    def __init__(self, model: nn.Module, inference_config=None, generation_config=None, pad_token_id: int = 0) -> None:
        from colossalai_b200.inference import InferenceConfig, InferenceEngine
        from colossalai_b200.inference.config import GenerationConfig

        self.engine = InferenceEngine(model, None, inference_config or InferenceConfig())
        self.gen_cfg = generation_config or GenerationConfig(max_new_tokens=64, do_sample=True)
        self.pad_token_id = pad_token_id

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, num_generations: int) -> torch.Tensor:
        prompts = []
        for row, m in zip(input_ids.tolist(), attention_mask.tolist()):
            p = [t for t, k in zip(row, m) if k]
            prompts += [p] * num_generations
        _, outs = self.engine.generate(prompts_token_ids=prompts, return_token_ids=True, generation_config=self.gen_cfg)
        P = input_ids.shape[1]
        A = max(len(o) - len(p) for o, p in zip(outs, prompts))
        rows = []
        for o, p in zip(outs, prompts):
            gen = o[len(p):]
            rows.append([self.pad_token_id] * (P - len(p)) + p + gen + [self.pad_token_id] * (A - len(gen)))
        return torch.tensor(rows, dtype=torch.long, device=input_ids.device)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.engine.model.load_state_dict(sd, strict=False)


class Producer:
    def __init__(self, backend, prompt_dataloader, num_generations: int = 4) -> None:
        self.backend, self.num_generations = backend, num_generations
        self._loader, self._it = prompt_dataloader, None
        self.model_version = 0

    def _next_prompts(self) -> Dict:
        if self._it is None:
            self._it = iter(self._loader)
        try:
            return next(self._it)
        except StopIteration:
            self._it = iter(self._loader)
            return next(self._it)

    def rollout(self) -> Dict:
        batch = self._next_prompts()
        ids, am = batch["input_ids"], batch.get("attention_mask")
        am = am if am is not None else torch.ones_like(ids)
        seq = self.backend.generate(ids, am, self.num_generations)
        out = {"sequences": seq, "prompt_len": ids.shape[1],
               "attention_mask": am.repeat_interleave(self.num_generations, 0), "model_version": self.model_version}
        for k, v in batch.items():
            if isinstance(v, list):
                out[k] = [x for x in v for _ in range(self.num_generations)]
        return out

    def sync_weights(self, state_dict: Dict[str, torch.Tensor], version: int) -> None:
        self.backend.load_state_dict(state_dict)
        self.model_version = version
