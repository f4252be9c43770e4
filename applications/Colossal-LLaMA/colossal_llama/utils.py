"""Vocabulary expansion, resumable checkpoints, streaming chat.
Parity: reference `colossal_llama/utils/{ckpt_io.py:36-99, stream_chat_patch.py, init_model.py (mean-init of new
embeddings), froze.py}`."""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Iterator, List, Optional, Tuple

import torch
import torch.nn as nn


@torch.no_grad()
def expand_vocab(model: nn.Module, new_vocab_size: int, new_token_sources: Optional[Dict[int, List[int]]] = None) -> nn.Module:
    """Grow the input embedding and LM head to `new_vocab_size`.  A new token is initialised with the MEAN of the old
    tokens it decomposes into (`new_token_sources[new_id] = [old ids]`), else with the mean of the whole table."""
    emb: nn.Embedding = model.model.embed_tokens
    head: nn.Linear = model.lm_head
    old = emb.weight.shape[0]
    assert new_vocab_size >= old
    tied = head.weight.data_ptr() == emb.weight.data_ptr()

    def grow(w: torch.Tensor) -> torch.Tensor:
        out = torch.empty(new_vocab_size, w.shape[1], dtype=w.dtype, device=w.device)
        out[:old] = w
        out[old:] = w.float().mean(0).to(w.dtype)
        for nid, src in (new_token_sources or {}).items():
            out[nid] = w[torch.tensor(src)].float().mean(0).to(w.dtype)
        return out

    new_emb = nn.Embedding(new_vocab_size, emb.weight.shape[1], padding_idx=emb.padding_idx)
    new_emb.weight = nn.Parameter(grow(emb.weight.data))
    model.model.embed_tokens = new_emb
    new_head = nn.Linear(head.in_features, new_vocab_size, bias=False)
    new_head.weight = new_emb.weight if tied else nn.Parameter(grow(head.weight.data))
    model.lm_head = new_head
    cfg = model.cfg.replace(vocab_size=new_vocab_size)
    model.cfg = model.config = cfg
    model.model.cfg = cfg
    return model


def save_checkpoint(save_dir: str, booster, model, optimizer, lr_scheduler, epoch: int, step: int, batch_size: int,
                    coordinator=None, sampler_start_idx: Optional[int] = None) -> str:
    """`save_dir/epoch-E_step-S/{modeling, optimizer, lr_scheduler, running_states.json}`."""
    path = os.path.join(save_dir, f"epoch-{epoch}_step-{step}")
    os.makedirs(os.path.join(path, "modeling"), exist_ok=True)
    booster.save_model(model, os.path.join(path, "modeling"), shard=True)
    booster.save_optimizer(optimizer, os.path.join(path, "optimizer"), shard=True)
    if lr_scheduler is not None:
        booster.save_lr_scheduler(lr_scheduler, os.path.join(path, "lr_scheduler"))
    state = {"epoch": epoch, "step": step,
             "sample_start_index": sampler_start_idx if sampler_start_idx is not None else step * batch_size}
    if coordinator is None or coordinator.is_master():
        with open(os.path.join(path, "running_states.json"), "w") as f:
            json.dump(state, f, indent=2)
    return path


def load_checkpoint(load_dir: str, booster, model, optimizer, lr_scheduler) -> Tuple[int, int, int]:
    booster.load_model(model, os.path.join(load_dir, "modeling"))
    booster.load_optimizer(optimizer, os.path.join(load_dir, "optimizer"))
    if lr_scheduler is not None and os.path.exists(os.path.join(load_dir, "lr_scheduler")):
        booster.load_lr_scheduler(lr_scheduler, os.path.join(load_dir, "lr_scheduler"))
    with open(os.path.join(load_dir, "running_states.json")) as f:
        s = json.load(f)
    return s["epoch"], s["step"], s["sample_start_index"]


@torch.no_grad()
def stream_chat(model: nn.Module, tokenizer, history: List[Dict[str, str]], query: str, max_new_tokens: int = 128,
                temperature: float = 0.7, top_p: float = 0.95, eos_token_id: int = 2) -> Iterator[Tuple[str, List[Dict]]]:
    """Yield (partial response, updated history) after every generated token."""
    text = "".join(f"<|{m['role']}|>\n{m['content']}\n" for m in history) + f"<|user|>\n{query}\n<|assistant|>\n"
    ids = torch.tensor([tokenizer(text)], device=next(model.parameters()).device)
    out: List[int] = []
    model.eval()
    for _ in range(max_new_tokens):
        logits = model(input_ids=ids)["logits"]
        logits = logits.reshape(1, ids.shape[1], -1)[0, -1, : model.cfg.vocab_size].float() / max(temperature, 1e-5)
        sl, si = logits.sort(descending=True)
        p = sl.softmax(-1)
        keep = (p.cumsum(-1) - p) <= top_p
        p = torch.where(keep, p, torch.zeros_like(p))
        nxt = int(si[torch.multinomial(p / p.sum(), 1)])
        if nxt == eos_token_id:
            break
        out.append(nxt)
        ids = torch.cat([ids, torch.tensor([[nxt]], device=ids.device)], 1)
        resp = tokenizer.decode(out) if hasattr(tokenizer, "decode") else str(out)
        yield resp, history + [{"role": "user", "content": query}, {"role": "assistant", "content": resp}]
