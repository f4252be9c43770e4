"""Batched model wrapper for the evaluation pipeline: per-choice loss scoring, masked target loss (perplexity-style
datasets) and greedy generation, each over right-padded batches, sharded over the data-parallel ranks.

Works with any causal LM that maps `input_ids [B, S]` (+ optional `attention_mask`) to logits (`{"logits": ...}` or
`.logits`, token-major or `[B, S, V]`): the native zoo, a `Booster`-wrapped model, a `transformers` model.  Right padding
is enough for a causal model - a row's real tokens never attend to the padding behind them - so generation keeps every
row at its own length and reads the logits at each row's last real token.

Parity: reference `colossal_eval/models/{base.py, huggingface.py:1-600}` (`HuggingFaceModel.inference`: `get_loss`
over `all_classes` / `calculate_loss` datasets, batched `generate`, first-token `logits_over_choices`), `models/vllm.py`.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

__all__ = ["EvalModel"]


class EvalModel:
    def __init__(self, model, tokenizer: Callable[[str], Sequence[int]], decode: Optional[Callable] = None,
                 batch_size: int = 8, max_new_tokens: int = 32, eos_token_id: int = 2, pad_token_id: int = 0,
                 bos_token_id: Optional[int] = None, max_length: int = 2048) -> None:
        self.model = model.eval()
        self.tokenizer, self.decode = tokenizer, decode
        self.batch_size, self.max_new_tokens, self.max_length = batch_size, max_new_tokens, max_length
        self.eos, self.pad, self.bos = eos_token_id, pad_token_id, bos_token_id
        self.device = next(model.parameters()).device

    # ------------------------------------------------------------------------------------------------ plumbing
    def _encode(self, text: str) -> List[int]:
        ids = list(self.tokenizer(text))
        return ([self.bos] if self.bos is not None else []) + ids

    def _logits(self, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        out = self.model(input_ids=ids, attention_mask=mask)
        logits = out["logits"] if isinstance(out, dict) else out.logits
        vocab = getattr(getattr(self.model, "cfg", None), "vocab_size", logits.shape[-1])
        return logits.reshape(ids.shape[0], ids.shape[1], -1)[..., :vocab].float()

    def _pad(self, rows: Sequence[Sequence[int]], extra: int = 0):
        width = max(len(r) for r in rows) + extra
        ids = torch.full((len(rows), width), self.pad, dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = torch.tensor(r, dtype=torch.long)
        lens = torch.tensor([len(r) for r in rows])
        mask = (torch.arange(width)[None] < lens[:, None]).long()
        return ids.to(self.device), mask.to(self.device), lens.to(self.device)

    # ------------------------------------------------------------------------------------------------ scoring
    @torch.no_grad()
    def get_loss(self, prompts: Sequence[str], targets: Sequence[str], reduce: str = "mean") -> List[float]:
        """Negative log-likelihood of `target` given `prompt`, per pair (`mean` over the target tokens or `sum`)."""
        out: List[float] = []
        for s in range(0, len(prompts), self.batch_size):
            rows, spans = [], []
            for p, t in zip(prompts[s:s + self.batch_size], targets[s:s + self.batch_size]):
                pi, ti = self._encode(p), list(self.tokenizer(t))
                pi = pi[-(self.max_length - len(ti)):] if len(pi) + len(ti) > self.max_length else pi
                rows.append(pi + ti)
                spans.append((len(pi), len(pi) + len(ti)))
            ids, mask, _ = self._pad(rows)
            lp = F.log_softmax(self._logits(ids, mask)[:, :-1], -1).gather(-1, ids[:, 1:, None]).squeeze(-1)
            for i, (a, b) in enumerate(spans):
                tok = -lp[i, a - 1:b - 1]
                out.append(float(tok.mean() if reduce == "mean" else tok.sum()))
        return out

    def score_choices(self, items: Sequence[Dict], length_normalize: bool = True) -> List[Dict]:
        """Multiple choice by loss: adds `output` (index of the cheapest choice) and `choice_losses` to every item."""
        flat_p, flat_t, owner = [], [], []
        for i, it in enumerate(items):
            for c in it["choices"]:
                flat_p.append(it["instruction"])
                flat_t.append(c)
                owner.append(i)
        losses = self.get_loss(flat_p, flat_t, "mean" if length_normalize else "sum")
        res = [dict(it, choice_losses=[]) for it in items]
        for o, l in zip(owner, losses):
            res[o]["choice_losses"].append(l)
        for r in res:
            r["output"] = min(range(len(r["choice_losses"])), key=r["choice_losses"].__getitem__)
        return res

    # ------------------------------------------------------------------------------------------------ generation
    @torch.no_grad()
    def generate(self, prompts: Sequence[str], max_new_tokens: Optional[int] = None) -> List[str]:
        n_new = max_new_tokens or self.max_new_tokens
        texts: List[str] = []
        for s in range(0, len(prompts), self.batch_size):
            rows = [self._encode(p)[-(self.max_length - n_new):] for p in prompts[s:s + self.batch_size]]
            ids, mask, lens = self._pad(rows, extra=n_new)
            start = lens.clone()
            done = torch.zeros(len(rows), dtype=torch.bool, device=self.device)
            ar = torch.arange(len(rows), device=self.device)
            for _ in range(n_new):
                width = int(lens.max())
                logits = self._logits(ids[:, :width], mask[:, :width])
                nxt = logits[ar, lens - 1].argmax(-1)
                nxt = torch.where(done, torch.full_like(nxt, self.pad), nxt)
                ids[ar, lens] = nxt
                mask[ar, lens] = (~done).long()
                done = done | (nxt == self.eos)
                lens = lens + (~done).long()
                if bool(done.all()):
                    break
            for i in range(len(rows)):
                gen = ids[i, int(start[i]):int(lens[i])].tolist()
                gen = [t for t in gen if t != self.eos]
                texts.append(self.decode(gen) if self.decode else " ".join(map(str, gen)))
        return texts

    # ------------------------------------------------------------------------------------------------ datasets
    def inference(self, items: Sequence[Dict], rank: int = 0, world_size: int = 1) -> List[Dict]:
        """Answer this rank's share (`items[rank::world_size]`) of a dataset.  Items with `choices` are scored by loss,
        items with `calculate_loss` get the loss of their `target`, all others are generated.  Every answered item
        carries its original position in `index` so the shares can be merged back in order."""
        mine = [dict(it, index=i) for i, it in enumerate(items) if i % world_size == rank]
        choice = [it for it in mine if "choices" in it]
        lossy = [it for it in mine if it.get("calculate_loss") and "choices" not in it]
        gen = [it for it in mine if "choices" not in it and not it.get("calculate_loss")]
        out = self.score_choices(choice) if choice else []
        if lossy:
            for it, l in zip(lossy, self.get_loss([it["instruction"] for it in lossy], [it["target"] for it in lossy])):
                out.append(dict(it, loss=l, num_target_tokens=len(list(self.tokenizer(it["target"])))))
        if gen:
            for it, text in zip(gen, self.generate([it["instruction"] for it in gen])):
                out.append(dict(it, output=text))
        return sorted(out, key=lambda it: it["index"])
