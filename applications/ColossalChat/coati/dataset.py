"""Conversation templating + tokenisation for SFT / preference / prompt-only / KTO data and the matching collators.
Parity: reference `coati/dataset/{conversation.py, tokenization_utils.py, loader.py, utils.py}`."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch
from torch.utils.data import Dataset

IGNORE_INDEX = -100


@dataclass
class Conversation:
    """`{role: text}` turns rendered with simple role tags; `end_of_assistant` closes every assistant turn so the loss
    mask can cover exactly the assistant spans."""

    system_message: str = ""
    roles: Sequence[str] = ("user", "assistant")
    role_tags: Dict[str, str] = field(default_factory=lambda: {"system": "<|system|>\n", "user": "<|user|>\n",
                                                               "assistant": "<|assistant|>\n"})
    end_of_assistant: str = "<|end|>\n"
    messages: List[Dict[str, str]] = field(default_factory=list)

    def append_message(self, role: str, content: str) -> None:
        assert role in self.roles or role == "system"
        self.messages.append({"role": role, "content": content})

    def clear(self) -> None:
        self.messages = []

    def segments(self, add_generation_prompt: bool = False) -> List[Dict[str, str]]:
        """[(text, is_assistant_content)] pieces in order."""
        segs = []
        if self.system_message:
            segs.append({"text": self.role_tags["system"] + self.system_message + "\n", "train": False})
        for m in self.messages:
            tag = self.role_tags[m["role"]]
            if m["role"] == "assistant":
                segs.append({"text": tag, "train": False})
                segs.append({"text": m["content"] + self.end_of_assistant, "train": True})
            else:
                segs.append({"text": tag + m["content"] + "\n", "train": False})
        if add_generation_prompt:
            segs.append({"text": self.role_tags["assistant"], "train": False})
        return segs

    def get_prompt(self, add_generation_prompt: bool = False) -> str:
        return "".join(s["text"] for s in self.segments(add_generation_prompt))


def _encode(tokenizer: Callable, text: str) -> List[int]:
    out = tokenizer(text)
    ids = out["input_ids"] if isinstance(out, dict) else out
    return list(ids[0]) if (len(ids) and isinstance(ids[0], (list, tuple))) else list(ids)


def tokenize_sft(messages: List[Dict[str, str]], tokenizer: Callable, conv: Optional[Conversation] = None,
                 max_length: int = 4096) -> Dict[str, List[int]]:
    """Whole conversation -> input_ids + labels where only assistant content (and its end tag) carries loss."""
    conv = conv or Conversation()
    conv.clear()
    for m in messages:
        conv.append_message(m["role"] if "role" in m else m["from"], m.get("content", m.get("value")))
    ids, labels = [], []
    for seg in conv.segments():
        t = _encode(tokenizer, seg["text"])
        ids += t
        labels += t if seg["train"] else [IGNORE_INDEX] * len(t)
    return {"input_ids": ids[:max_length], "labels": labels[:max_length]}


def tokenize_prompt(messages: List[Dict[str, str]], tokenizer: Callable, conv: Optional[Conversation] = None,
                    max_length: int = 4096) -> Dict[str, List[int]]:
    conv = conv or Conversation()
    conv.clear()
    for m in messages:
        conv.append_message(m["role"], m["content"])
    return {"input_ids": _encode(tokenizer, conv.get_prompt(add_generation_prompt=True))[-max_length:]}


def tokenize_preference(context: List[Dict[str, str]], chosen: str, rejected: str, tokenizer: Callable,
                        conv: Optional[Conversation] = None, max_length: int = 4096) -> Dict[str, List[int]]:
    """Shared prompt + two completions; `*_loss_mask` marks the completion tokens."""
    p = tokenize_prompt(context, tokenizer, conv, max_length)["input_ids"]
    end = (conv or Conversation()).end_of_assistant
    out = {}
    for name, text in (("chosen", chosen), ("rejected", rejected)):
        c = _encode(tokenizer, text + end)
        ids = (p + c)[:max_length]
        out[f"{name}_input_ids"] = ids
        out[f"{name}_loss_mask"] = ([0] * len(p) + [1] * len(c))[:max_length]
    return out


def tokenize_kto(context: List[Dict[str, str]], completion: str, label: bool, tokenizer: Callable,
                 conv: Optional[Conversation] = None, max_length: int = 4096) -> Dict:
    p = tokenize_prompt(context, tokenizer, conv, max_length)["input_ids"]
    c = _encode(tokenizer, completion + (conv or Conversation()).end_of_assistant)
    return {"input_ids": (p + c)[:max_length], "loss_mask": ([0] * len(p) + [1] * len(c))[:max_length], "label": label,
            "prompt_len": len(p)}


class ListDataset(Dataset):
    def __init__(self, items: List[dict]) -> None:
        self.items = items

    def __len__(self) -> int:
        return len(self.items)

    def __getitem__(self, i: int) -> dict:
        return self.items[i]


def _pad(seqs: List[List[int]], value: int, left: bool = False, max_length: Optional[int] = None) -> torch.Tensor:
    n = max_length or max(len(s) for s in seqs)
    rows = []
    for s in seqs:
        s = list(s)[:n]
        pad = [value] * (n - len(s))
        rows.append(pad + s if left else s + pad)
    return torch.tensor(rows, dtype=torch.long)


@dataclass
class DataCollatorForSupervisedDataset:
    pad_token_id: int = 0
    max_length: Optional[int] = None

    def __call__(self, batch: List[dict]) -> Dict[str, torch.Tensor]:
        ids = _pad([b["input_ids"] for b in batch], self.pad_token_id, max_length=self.max_length)
        labels = _pad([b["labels"] for b in batch], IGNORE_INDEX, max_length=self.max_length)
        mask = _pad([[1] * len(b["input_ids"]) for b in batch], 0, max_length=self.max_length)
        return {"input_ids": ids, "labels": labels, "attention_mask": mask}


@dataclass
class DataCollatorForPromptDataset:
    pad_token_id: int = 0

    def __call__(self, batch: List[dict]) -> Dict[str, torch.Tensor]:
        ids = _pad([b["input_ids"] for b in batch], self.pad_token_id, left=True)      # left padding for generation
        mask = _pad([[1] * len(b["input_ids"]) for b in batch], 0, left=True)
        out = {"input_ids": ids, "attention_mask": mask}
        if "gt_answer" in batch[0]:
            out["gt_answer"] = [b["gt_answer"] for b in batch]
        return out


@dataclass
class DataCollatorForPreferenceDataset:
    pad_token_id: int = 0

    def __call__(self, batch: List[dict]) -> Dict[str, torch.Tensor]:
        out = {}
        for name in ("chosen", "rejected"):
            out[f"{name}_input_ids"] = _pad([b[f"{name}_input_ids"] for b in batch], self.pad_token_id)
            out[f"{name}_loss_mask"] = _pad([b[f"{name}_loss_mask"] for b in batch], 0)
            out[f"{name}_attention_mask"] = _pad([[1] * len(b[f"{name}_input_ids"]) for b in batch], 0)
        return out


@dataclass
class DataCollatorForKTODataset:
    pad_token_id: int = 0

    def __call__(self, batch: List[dict]) -> Dict[str, torch.Tensor]:
        ids = _pad([b["input_ids"] for b in batch], self.pad_token_id)
        # KL samples: every prompt paired with the completion of the NEXT sample (mismatched pairs estimate z0)
        n = len(batch)
        kl_ids, kl_mask = [], []
        for i, b in enumerate(batch):
            o = batch[(i + 1) % n]
            comp = o["input_ids"][o["prompt_len"]:]
            kl_ids.append(b["input_ids"][: b["prompt_len"]] + comp)
            kl_mask.append([0] * b["prompt_len"] + [1] * len(comp))
        return {"input_ids": ids, "loss_mask": _pad([b["loss_mask"] for b in batch], 0),
                "attention_mask": _pad([[1] * len(b["input_ids"]) for b in batch], 0),
                "label": torch.tensor([bool(b["label"]) for b in batch]),
                "kl_input_ids": _pad(kl_ids, self.pad_token_id), "kl_loss_mask": _pad(kl_mask, 0),
                "kl_attention_mask": _pad([[1] * len(k) for k in kl_ids], 0)}
