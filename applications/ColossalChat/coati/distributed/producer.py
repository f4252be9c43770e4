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
    """`producer_idx` / `num_producers`: several producers draw from the same prompt dataloader and each keeps every
    `num_producers`-th batch (the reference gives every producer a `DistributedSampler` shard, `producer.py:120-150`).
    `rollout_log`: a jsonl file that receives one line per rollout batch (token ids, or text when a `tokenizer` with
    `batch_decode` is given) - the reference's `rollout_log_file`.  `evaluate` scores greedy / sampled completions of
    held-out prompts with a verifiable reward (the reference's periodic eval over `eval_dataset_config`)."""

    def __init__(self, backend, prompt_dataloader, num_generations: int = 4, producer_idx: int = 0,
                 num_producers: int = 1, rollout_log: Optional[str] = None, tokenizer=None) -> None:
        assert 0 <= producer_idx < num_producers
        self.backend, self.num_generations = backend, num_generations
        self.producer_idx, self.num_producers = producer_idx, num_producers
        self._loader, self._it, self._drawn = prompt_dataloader, None, 0
        self.rollout_log, self.tokenizer = rollout_log, tokenizer
        self.model_version = 0

    def _draw(self) -> Dict:
        if self._it is None:
            self._it = iter(self._loader)
        try:
            return next(self._it)
        except StopIteration:
            self._it = iter(self._loader)
            return next(self._it)

    def _next_prompts(self) -> Dict:
        while True:
            batch = self._draw()
            mine = self._drawn % self.num_producers == self.producer_idx
            self._drawn += 1
            if mine:
                return batch

    def _log(self, rollout: Dict) -> None:
        import json

        seq, P = rollout["sequences"], rollout["prompt_len"]
        if self.tokenizer is not None:
            rec = {"prompt": self.tokenizer.batch_decode(seq[:, :P].tolist()),
                   "response": self.tokenizer.batch_decode(seq[:, P:].tolist())}
        else:
            rec = {"prompt_ids": seq[:, :P].tolist(), "response_ids": seq[:, P:].tolist()}
        rec.update(model_version=self.model_version, producer=self.producer_idx)
        with open(self.rollout_log, "a") as f:
            f.write(json.dumps(rec) + "\n")

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
        if self.rollout_log is not None:
            self._log(out)
        return out

    @torch.no_grad()
    def evaluate(self, eval_dataloaders: Dict[str, object], reward_fn, num_generations: int = 1,
                 max_batches: Optional[int] = None) -> Dict[str, float]:
        """Mean reward of `num_generations` completions per held-out prompt, per dataset: {"eval/<name>": score}."""
        scores: Dict[str, float] = {}
        for name, loader in eval_dataloaders.items():
            total, n = 0.0, 0
            for i, batch in enumerate(loader):
                if max_batches is not None and i >= max_batches:
                    break
                ids, am = batch["input_ids"], batch.get("attention_mask")
                am = am if am is not None else torch.ones_like(ids)
                seq = self.backend.generate(ids, am, num_generations)
                extra = {k: [x for x in v for _ in range(num_generations)] for k, v in batch.items()
                         if isinstance(v, list)}
                r = reward_fn(seq, ids.shape[1], **extra).float()
                total, n = total + float(r.sum()), n + r.numel()
            scores[f"eval/{name}"] = total / max(n, 1)
        return scores

    def sync_weights(self, state_dict: Dict[str, torch.Tensor], version: int) -> None:
        self.backend.load_state_dict(state_dict)
        self.model_version = version
