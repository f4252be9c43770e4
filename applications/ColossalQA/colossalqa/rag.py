"""Parity: reference `colossalqa/{text_splitter/, retriever.py (CustomRetriever with per-source indexes + sql record
manager), memory.py (ConversationBufferWithSummary), chain/retrieval_qa/, prompt/prompt.py}` — the langchain plumbing
is replaced by small explicit classes; any embedding function and any `generate(prompt) -> str` callable plug in
(e.g. our `InferenceEngine`)."""
from __future__ import annotations

import hashlib
import math
import re
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch


def split_text(text: str, chunk_size: int = 256, chunk_overlap: int = 32) -> List[str]:
    """Sentence-aware splitter: sentences are packed into chunks of at most `chunk_size` characters, consecutive
    chunks share roughly `chunk_overlap` characters of context."""
    sents = [s.strip() for s in re.split(r"(?<=[.!?。！？])\s+|\n+", text) if s.strip()]
    chunks, cur = [], ""
    for s in sents:
        if cur and len(cur) + 1 + len(s) > chunk_size:
            chunks.append(cur)
            cur = cur[-chunk_overlap:].lstrip() if chunk_overlap else ""
        cur = (cur + " " + s).strip() if cur else s
        while len(cur) > chunk_size:                      # a single over-long sentence
            chunks.append(cur[:chunk_size])
            cur = cur[chunk_size - chunk_overlap:]
    if cur:
        chunks.append(cur)
    return chunks


def hashing_embedder(dim: int = 256) -> Callable[[Sequence[str]], torch.Tensor]:
    """Dependency-free bag-of-words embedder (feature hashing + l2 norm) — the default when no model is supplied."""

    def embed(texts: Sequence[str]) -> torch.Tensor:
        out = torch.zeros(len(texts), dim)
        for i, t in enumerate(texts):
            for w in re.findall(r"\w+", t.lower()):
                h = int(hashlib.md5(w.encode()).hexdigest(), 16)
                out[i, h % dim] += 1.0 if (h >> 20) & 1 else -1.0
        return torch.nn.functional.normalize(out, dim=-1)

    return embed


class EmbeddingIndex:
    def __init__(self, embed: Optional[Callable[[Sequence[str]], torch.Tensor]] = None) -> None:
        self.embed = embed or hashing_embedder()
        self.docs: List[Dict] = []
        self.matrix: Optional[torch.Tensor] = None
        self._seen: set = set()

    def add_documents(self, texts: Sequence[str], source: str = "default", cleanup: str = "incremental") -> int:
        """Deduplicated insert (content hash), like the reference's record manager in `incremental` mode."""
        new = []
        for t in texts:
            h = hashlib.sha1((source + "\0" + t).encode()).hexdigest()
            if h not in self._seen:
                self._seen.add(h)
                new.append({"text": t, "source": source})
        if new:
            vec = self.embed([d["text"] for d in new])
            self.docs += new
            self.matrix = vec if self.matrix is None else torch.cat([self.matrix, vec], 0)
        return len(new)

    def search(self, query: str, k: int = 3, source: Optional[str] = None) -> List[Tuple[Dict, float]]:
        if self.matrix is None:
            return []
        sims = (self.matrix @ self.embed([query])[0]).tolist()
        idx = [i for i in range(len(sims)) if source is None or self.docs[i]["source"] == source]
        idx.sort(key=lambda i: -sims[i])
        return [(self.docs[i], sims[i]) for i in idx[:k]]


class ConversationMemory:
    """Keeps the last turns verbatim; older turns are folded into a running summary by `summarize` (any callable)."""

    def __init__(self, max_turns: int = 4, summarize: Optional[Callable[[str], str]] = None) -> None:
        self.max_turns, self.summarize = max_turns, summarize
        self.turns: List[Tuple[str, str]] = []
        self.summary = ""

    def add(self, user: str, assistant: str) -> None:
        self.turns.append((user, assistant))
        if len(self.turns) > self.max_turns:
            old = self.turns.pop(0)
            text = (self.summary + f"\nuser: {old[0]}\nassistant: {old[1]}").strip()
            self.summary = self.summarize(text) if self.summarize else text[-512:]

    def render(self) -> str:
        hist = "\n".join(f"user: {u}\nassistant: {a}" for u, a in self.turns)
        return (f"summary: {self.summary}\n" if self.summary else "") + hist


class RetrievalQA:
    PROMPT = ("Answer the question using only the context. If the context is not enough, say you do not know.\n"
              "{memory}\ncontext:\n{context}\nquestion: {question}\nanswer:")

    def __init__(self, index, generate: Callable[[str], str], k: int = 3, min_score: float = 0.05,
                 memory: Optional[ConversationMemory] = None, rewrite: bool = False) -> None:
        """`index`: anything with `search(query, k, source=None)` (EmbeddingIndex, BM25Index, HybridRetriever).
        `rewrite`: make follow-up questions stand alone (with the same generator) before retrieving."""
        self.index, self.generate, self.k, self.min_score = index, generate, k, min_score
        self.memory = memory or ConversationMemory()
        self.rewrite = rewrite

    def build_prompt(self, question: str) -> Tuple[str, List[Dict]]:
        query = question
        if self.rewrite:
            from .retrieval import rewrite_follow_up

            query = rewrite_follow_up(question, self.memory, self.generate)
        hits = [(d, s) for d, s in self.index.search(query, self.k) if s >= self.min_score]
        ctx = "\n".join(f"[{i + 1}] ({d['source']}) {d['text']}" for i, (d, _) in enumerate(hits))
        return self.PROMPT.format(memory=self.memory.render(), context=ctx or "(none)", question=question), [d for d, _ in hits]

    def run(self, question: str) -> Tuple[str, List[Dict]]:
        prompt, sources = self.build_prompt(question)
        answer = self.generate(prompt) if sources else "I do not know."
        self.memory.add(question, answer)
        return answer, sources
