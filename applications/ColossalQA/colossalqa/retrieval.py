"""Lexical + hybrid retrieval, query rewriting and document loading (reference `colossalqa/retriever.py`
`CustomRetriever` with per-source indices and k-per-source merging, `colossalqa/data_loader/document_loader.py`,
`colossalqa/chain/retrieval_qa/base.py` follow-up-question disambiguation).

* `BM25Index`     - Okapi BM25 over whitespace / CJK-character tokens, per-source filtering, incremental adds;
* `HybridRetriever` - reciprocal-rank fusion of any number of indices exposing `search(query, k, source)`
                    (typically `EmbeddingIndex` + `BM25Index`);
* `rewrite_follow_up` - makes a follow-up question stand alone using the conversation memory and any generator callable;
* `load_documents` - .txt / .md (split on headings first) / .jsonl / .csv files -> chunks with `source` metadata."""
from __future__ import annotations

import csv
import json
import math
import re
from collections import Counter
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

from .rag import split_text

__all__ = ["tokenize", "BM25Index", "HybridRetriever", "rewrite_follow_up", "load_documents"]

_CJK = re.compile(r"[一-鿿぀-ヿ가-힯]")


def tokenize(text: str) -> List[str]:
    """Lower-cased word tokens; CJK characters are tokens of their own (no word segmenter needed)."""
    out: List[str] = []
    for piece in re.findall(r"[A-Za-z0-9_]+|[^\sA-Za-z0-9_]", text.lower()):
        if piece.isalnum() or "_" in piece or _CJK.match(piece):
            out.append(piece)
    return out


class BM25Index:
    def __init__(self, k1: float = 1.5, b: float = 0.75) -> None:
        self.k1, self.b = k1, b
        self.docs: List[Dict] = []
        self._tf: List[Counter] = []
        self._len: List[int] = []
        self._df: Counter = Counter()
        self._seen: set = set()

    def add_documents(self, texts: Sequence[str], source: str = "default") -> int:
        n = 0
        for t in texts:
            key = (source, t)
            if key in self._seen:
                continue
            self._seen.add(key)
            toks = tokenize(t)
            tf = Counter(toks)
            self.docs.append({"text": t, "source": source})
            self._tf.append(tf)
            self._len.append(len(toks))
            self._df.update(tf.keys())
            n += 1
        return n

    def search(self, query: str, k: int = 3, source: Optional[str] = None) -> List[Tuple[Dict, float]]:
        if not self.docs:
            return []
        N = len(self.docs)
        avg = sum(self._len) / N
        q = tokenize(query)
        scored = []
        for i, tf in enumerate(self._tf):
            if source is not None and self.docs[i]["source"] != source:
                continue
            s = 0.0
            for w in q:
                f = tf.get(w, 0)
                if not f:
                    continue
                idf = math.log(1.0 + (N - self._df[w] + 0.5) / (self._df[w] + 0.5))
                s += idf * f * (self.k1 + 1) / (f + self.k1 * (1 - self.b + self.b * self._len[i] / max(avg, 1e-9)))
            if s > 0:
                scored.append((i, s))
        scored.sort(key=lambda x: -x[1])
        return [(self.docs[i], s) for i, s in scored[:k]]


class HybridRetriever:
    """Reciprocal-rank fusion: score(d) = sum over retrievers of weight / (rrf_k + rank).  Exposes the same `search`
    signature as the single indices, so `RetrievalQA` takes it unchanged; `sources` restricts / balances the result to
    at most `k_per_source` chunks per source (the reference retriever's per-index top-k)."""

    def __init__(self, retrievers: Sequence, weights: Optional[Sequence[float]] = None, rrf_k: int = 60,
                 fetch_k: int = 10, k_per_source: Optional[int] = None) -> None:
        self.retrievers = list(retrievers)
        self.weights = list(weights) if weights is not None else [1.0] * len(self.retrievers)
        self.rrf_k, self.fetch_k, self.k_per_source = rrf_k, fetch_k, k_per_source

    def add_documents(self, texts: Sequence[str], source: str = "default") -> int:
        return max(r.add_documents(texts, source=source) for r in self.retrievers)

    def search(self, query: str, k: int = 3, source: Optional[str] = None) -> List[Tuple[Dict, float]]:
        fused: Dict[Tuple[str, str], float] = {}
        docs: Dict[Tuple[str, str], Dict] = {}
        for r, w in zip(self.retrievers, self.weights):
            for rank, (d, _) in enumerate(r.search(query, self.fetch_k, source)):
                key = (d["source"], d["text"])
                docs[key] = d
                fused[key] = fused.get(key, 0.0) + w / (self.rrf_k + rank + 1)
        ranked = sorted(fused.items(), key=lambda kv: -kv[1])
        out, per = [], Counter()
        for key, s in ranked:
            if self.k_per_source is not None and per[key[0]] >= self.k_per_source:
                continue
            per[key[0]] += 1
            out.append((docs[key], s * (self.rrf_k + 1)))      # 1.0 = ranked first by one retriever of weight 1
            if len(out) == k:
                break
        return out


_REWRITE = ("Rewrite the follow-up question so that it can be understood without the conversation. Keep it one "
            "sentence; do not answer it.\n{history}\nfollow-up: {question}\nstand-alone question:")


def rewrite_follow_up(question: str, memory, generate: Callable[[str], str]) -> str:
    """Pronoun / ellipsis resolution before retrieval; returns the question unchanged when there is no history or the
    generator gives nothing usable."""
    history = memory.render() if memory is not None else ""
    if not history.strip():
        return question
    out = (generate(_REWRITE.format(history=history, question=question)) or "").strip().splitlines()
    out = out[0].strip() if out else ""
    return out if 3 <= len(out) <= 4 * max(len(question), 40) else question


def load_documents(paths: Sequence, chunk_size: int = 256, chunk_overlap: int = 32,
                   text_key: str = "text") -> List[Dict]:
    """[{ "text", "source" }] chunks.  Markdown is split on headings before the length-based splitter so that a chunk
    never straddles two sections; jsonl / csv rows are one document each (`text_key` column / field)."""
    out: List[Dict] = []
    for p in map(Path, paths):
        suffix = p.suffix.lower()
        if suffix in (".txt", ""):
            pieces = [p.read_text()]
        elif suffix in (".md", ".markdown"):
            pieces = [s for s in re.split(r"(?m)^(?=#{1,6} )", p.read_text()) if s.strip()]
        elif suffix == ".jsonl":
            pieces = [json.loads(l)[text_key] for l in p.read_text().splitlines() if l.strip()]
        elif suffix == ".csv":
            with p.open(newline="") as f:
                pieces = [row[text_key] for row in csv.DictReader(f)]
        else:
            raise ValueError(f"unsupported document type: {p.name}")
        for piece in pieces:
            for c in split_text(piece, chunk_size, chunk_overlap):
                out.append({"text": c, "source": p.name})
    return out
