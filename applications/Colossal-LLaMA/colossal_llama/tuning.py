"""Fine-tuning helpers of the continual-pre-training recipe (reference `colossal_llama/utils/{neftune_patch.py,
froze.py, utils.py}`, `colossal_llama/tokenizer/init_tokenizer.py`).

* `activate_neftune` / `deactivate_neftune` - NEFTune: uniform noise of magnitude alpha / sqrt(L * d) added to the input
  embeddings while training (a forward hook on the embedding; works on sharded / wrapped models because it only sees
  the embedding output);
* `freeze_non_embeds_parameters` - stage 1 of vocabulary expansion: train the (new) embedding rows and LM head only;
* `unfreeze_parameters`;
* `plan_vocab_expansion` - given the old tokenizer's encoding of every candidate token, decide which tokens to add
  (frequency x compression gain) and return the `new_token_sources` map that `expand_vocab` consumes;
* `format_numel_str`, `get_model_numel`."""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

__all__ = ["activate_neftune", "deactivate_neftune", "freeze_non_embeds_parameters", "unfreeze_parameters",
           "plan_vocab_expansion", "get_model_numel", "format_numel_str"]


def _embedding_of(model: nn.Module) -> nn.Module:
    inner = model.unwrap() if hasattr(model, "unwrap") else model
    for name in ("model.embed_tokens", "embed_tokens", "transformer.wte", "wte"):
        mod = inner
        try:
            for part in name.split("."):
                mod = getattr(mod, part)
            if mod is not None:
                return mod
        except AttributeError:
            continue
    for mod in inner.modules():
        if isinstance(mod, nn.Embedding):
            return mod
    raise ValueError("no input embedding found")


def activate_neftune(model: nn.Module, neftune_noise_alpha: float = 5.0, generator: Optional[torch.Generator] = None):
    """Returns the model; the hook handle is kept on the embedding (`_neftune_handle`) for `deactivate_neftune`."""
    emb = _embedding_of(model)
    deactivate_neftune(model)

    def hook(module, inputs, output):
        if not module.training:
            return output
        dims = output.shape[-2] * output.shape[-1] if output.dim() >= 3 else output.shape[0] * output.shape[-1]
        mag = neftune_noise_alpha / math.sqrt(dims)
        noise = torch.empty(output.shape, dtype=torch.float32, device=output.device).uniform_(-mag, mag, generator=generator)
        return output + noise.to(output.dtype)

    emb._neftune_handle = emb.register_forward_hook(hook)
    emb.neftune_noise_alpha = neftune_noise_alpha
    return model


def deactivate_neftune(model: nn.Module):
    emb = _embedding_of(model)
    handle = getattr(emb, "_neftune_handle", None)
    if handle is not None:
        handle.remove()
        emb._neftune_handle = None
    return model


def freeze_non_embeds_parameters(model: nn.Module, also_train: Sequence[str] = ("lm_head",)) -> List[str]:
    """Everything but the input embedding (and modules whose name contains one of `also_train`) stops requiring grad;
    returns the names that stay trainable."""
    emb = _embedding_of(model)
    emb_ids = {id(p) for p in emb.parameters()}
    keep = []
    for name, p in model.named_parameters():
        train = id(p) in emb_ids or any(k in name for k in also_train)
        p.requires_grad_(train)
        if train:
            keep.append(name)
    return keep


def unfreeze_parameters(model: nn.Module) -> None:
    for p in model.parameters():
        p.requires_grad_(True)


def plan_vocab_expansion(candidates: Dict[str, int], encode: Callable[[str], Sequence[int]], old_vocab_size: int,
                         max_new_tokens: int, min_pieces: int = 2) -> Tuple[List[str], Dict[int, List[int]]]:
    """candidates: {token string: corpus frequency}.  A candidate is worth adding when the old tokenizer needs at least
    `min_pieces` ids for it; the gain of adding it is frequency x (pieces - 1) ids saved.  Returns the chosen strings
    (new id = old_vocab_size + position) and the `{new id: old ids}` map for mean-initialising their embeddings."""
    scored = []
    for tok, freq in candidates.items():
        pieces = list(encode(tok))
        if len(pieces) >= min_pieces:
            scored.append((freq * (len(pieces) - 1), tok, pieces))
    scored.sort(key=lambda x: (-x[0], x[1]))
    chosen = scored[:max_new_tokens]
    tokens = [t for _, t, _ in chosen]
    sources = {old_vocab_size + i: p for i, (_, _, p) in enumerate(chosen)}
    return tokens, sources


def get_model_numel(model: nn.Module, trainable_only: bool = False) -> int:
    seen, n = set(), 0
    for p in model.parameters():
        if id(p) in seen or (trainable_only and not p.requires_grad):
            continue
        seen.add(id(p))
        n += p.numel()
    return n


def format_numel_str(numel: int) -> str:
    for unit, div in (("B", 1e9), ("M", 1e6), ("K", 1e3)):
        if numel >= div:
            return f"{numel / div:.2f} {unit}"
    return str(numel)
