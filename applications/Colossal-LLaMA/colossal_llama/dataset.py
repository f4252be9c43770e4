"""Pre-training data: tokenise documents (source kept out of the loss, target trained) and SPLICE them into
sequences of exactly `max_length` tokens so no compute is spent on padding.
Parity: reference `colossal_llama/dataset/spliced_and_tokenized_dataset.py:1-300`."""
from __future__ import annotations

import random
from typing import Callable, Dict, Iterable, Iterator, List, Optional

import torch
from torch.utils.data import IterableDataset

IGNORE_INDEX = -100


def supervised_tokenize_pretrain(data_point: Dict[str, str], tokenizer: Callable, bos: int = 1, eos: int = 2,
                                 ignore_index: int = IGNORE_INDEX, max_length: int = 4096) -> Dict[str, List[int]]:
    """`{"source": optional context, "target": text}` -> ids/labels (`source` tokens carry no loss)."""
    src = list(tokenizer(data_point.get("source", "") or ""))
    tgt = list(tokenizer(data_point["target"]))
    ids = ([bos] + src + tgt + [eos])[:max_length]
    labels = ([ignore_index] * (1 + len(src)) + tgt + [eos])[:max_length]
    return {"input_ids": ids, "labels": labels, "seq_length": len(ids)}


class ClosedToConstantLengthSplicedDataset(IterableDataset):
    """Greedy bin packing of tokenised samples into `max_length` slots out of a look-ahead buffer; every yielded item
    is padded only by the (small) remainder the buffer could not fill, and carries `seq_boundaries` (cu_seqlens) so
    attention can stay block-diagonal."""

    def __init__(self, dataset: Iterable[Dict[str, List[int]]], max_length: int = 4096, num_packed_sequences: int = 8,
                 pad_token_id: int = 0, shuffle: bool = False, seed: int = 0, infinite: bool = False,
                 error_strict: bool = False) -> None:
        self.dataset, self.max_length, self.buf_n = dataset, max_length, num_packed_sequences
        self.pad_token_id, self.shuffle, self.seed, self.infinite, self.error_strict = pad_token_id, shuffle, seed, infinite, error_strict
        self.current_size = 0

    def _emit(self, group: List[Dict[str, List[int]]]) -> Dict[str, torch.Tensor]:
        ids, labels, bounds = [], [], [0]
        for s in group:
            ids += s["input_ids"]
            labels += s["labels"]
            bounds.append(len(ids))
        pad = self.max_length - len(ids)
        self.current_size += 1
        return {"input_ids": torch.tensor(ids + [self.pad_token_id] * pad),
                "labels": torch.tensor(labels + [IGNORE_INDEX] * pad),
                "attention_mask": torch.tensor([1] * len(ids) + [0] * pad),
                "seq_boundaries": torch.tensor(bounds)}

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        rng = random.Random(self.seed)
        while True:
            it = iter(self.dataset)
            buf: List[Dict[str, List[int]]] = []
            exhausted = False
            while not exhausted or buf:
                while not exhausted and len(buf) < self.buf_n * 4:
                    try:
                        s = next(it)
                    except StopIteration:
                        exhausted = True
                        break
                    if len(s["input_ids"]) > self.max_length:
                        if self.error_strict:
                            raise ValueError(f"sample of {len(s['input_ids'])} tokens exceeds max_length")
                        s = {k: (v[: self.max_length] if isinstance(v, list) else v) for k, v in s.items()}
                    buf.append(s)
                if not buf:
                    break
                if self.shuffle:
                    rng.shuffle(buf)
                buf.sort(key=lambda s: -len(s["input_ids"]))         # first-fit decreasing
                group, room, rest = [], self.max_length, []
                for s in buf:
                    if len(s["input_ids"]) <= room and len(group) < self.buf_n:
                        group.append(s)
                        room -= len(s["input_ids"])
                    else:
                        rest.append(s)
                buf = rest
                yield self._emit(group)
            if not self.infinite:
                return
