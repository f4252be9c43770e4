"""Language-universal retrieval conversation + a local LLM wrapper + a table loader.

`UniversalRetrievalConversation` keeps one retriever and one memory per language (English / Chinese), detects the
language of every question from its CJK share, splits Chinese text on its own punctuation (no spaces to rely on),
uses per-language prompt templates, refuses (instead of hallucinating) when nothing relevant is retrieved, and can first
classify the user's INTENT (e.g. customer service: refund / shipping / other) to route a question to a canned flow.

`LocalLLM` turns any text generator - `colossalai_b200.inference.InferenceEngine`, ColossalEval's `EvalModel`, an HTTP
`/generate` endpoint - into the `prompt -> answer` callable the chains take, with stop words and a length cap.

Parity: reference `colossalqa/{retrieval_conversation_universal.py, retrieval_conversation_en.py,
retrieval_conversation_zh.py, local/llm.py (ColossalAPI / ColossalLLM), prompt/prompt.py,
text_splitter/chinese_text_splitter.py, data_loader/table_dataloader.py}` and
`examples/retrieval_intent_classification_zh_customer_service.py`.
"""
from __future__ import annotations

import csv
import json
import re
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

from .rag import ConversationMemory, RetrievalQA, split_text
from .retrieval import BM25Index

__all__ = ["detect_language", "split_chinese_text", "LocalLLM", "load_table", "UniversalRetrievalConversation",
           "classify_intent", "PROMPTS"]

PROMPTS = {
    "en": ("Answer the question using only the context. If the context is not enough, say you do not know.\n"
           "{memory}\ncontext:\n{context}\nquestion: {question}\nanswer:"),
    "zh": ("请只根据下面的资料回答问题；如果资料不足以回答，请回答“我不知道”。\n"
           "{memory}\n资料：\n{context}\n问题：{question}\n回答："),
}
REFUSAL = {"en": "I do not know.", "zh": "我不知道。"}


def detect_language(text: str) -> str:
    cjk = sum(1 for ch in text if "一" <= ch <= "鿿")
    letters = sum(1 for ch in text if ch.isalpha())
    return "zh" if letters and cjk / letters > 0.3 else "en"


def split_chinese_text(text: str, chunk_size: int = 128) -> List[str]:
    """Sentence-aware splitter for text without spaces: cut after 。！？；… (and newlines), then pack sentences into
    chunks of at most `chunk_size` characters; an over-long sentence is cut at commas, then hard."""
    sentences = [s for s in re.split(r"(?<=[。！？；…\n])", text) if s.strip()]
    pieces: List[str] = []
    for s in sentences:
        if len(s) <= chunk_size:
            pieces.append(s)
            continue
        for part in re.split(r"(?<=[，、,])", s):
            while len(part) > chunk_size:
                pieces.append(part[:chunk_size])
                part = part[chunk_size:]
            if part:
                pieces.append(part)
    chunks, cur = [], ""
    for p in pieces:
        if cur and len(cur) + len(p) > chunk_size:
            chunks.append(cur.strip())
            cur = ""
        cur += p
    if cur.strip():
        chunks.append(cur.strip())
    return chunks


class LocalLLM:
    def __init__(self, generate: Callable[..., object], stop: Sequence[str] = ("\nquestion:", "\n问题：", "\n\n"),
                 max_chars: int = 1024) -> None:
        """`generate(prompt) -> str | list[str]` (an engine's `generate(prompts=[...])` result is unwrapped)."""
        self._generate, self.stop, self.max_chars = generate, tuple(stop), max_chars
        self.calls = 0

    def __call__(self, prompt: str) -> str:
        self.calls += 1
        out = self._generate(prompt)
        if isinstance(out, (list, tuple)):
            out = out[0] if out else ""
        text = str(out)
        if text.startswith(prompt):                     # engines that echo the prompt
            text = text[len(prompt):]
        cut = min([i for i in (text.find(s) for s in self.stop) if i > 0] or [len(text)])
        return text[:cut].strip()[: self.max_chars]

    @classmethod
    def from_engine(cls, engine, **kw) -> "LocalLLM":
        return cls(lambda p: engine.generate(prompts=[p]), **kw)

    @classmethod
    def from_http(cls, url: str, max_new_tokens: int = 128, **kw) -> "LocalLLM":
        def call(p):
            import requests

            return requests.post(url.rstrip("/") + "/generate", json={"prompt": p, "max_new_tokens": max_new_tokens},
                                 timeout=120).json().get("text", "")
        return cls(call, **kw)


def load_table(path, columns: Optional[Sequence[str]] = None, source: Optional[str] = None) -> List[Dict]:
    """One document per row of a csv / jsonl / json-array table: "col: value; col: value; ..." (what a retriever can
    match on) with the raw row kept under `row`."""
    p = Path(path)
    if p.suffix.lower() == ".csv":
        with p.open(newline="") as f:
            rows = list(csv.DictReader(f))
    elif p.suffix.lower() == ".jsonl":
        rows = [json.loads(l) for l in p.read_text().splitlines() if l.strip()]
    else:
        rows = json.loads(p.read_text())
    docs = []
    for r in rows:
        keys = list(columns) if columns else list(r)
        docs.append({"text": "; ".join(f"{k}: {r[k]}" for k in keys if r.get(k) not in (None, "")),
                     "source": source or p.name, "row": r})
    return docs


def classify_intent(llm: Callable[[str], str], question: str, intents: Dict[str, str], default: str = "other") -> str:
    """Ask the model to pick one intent label; anything that is not exactly a known label falls back to `default`."""
    menu = "\n".join(f"- {k}: {v}" for k, v in intents.items())
    reply = llm(f"Classify the user's request into exactly one of these intents and answer with the label only.\n{menu}\n"
                f"request: {question}\nintent:")
    reply = reply.strip().lower().strip(".:\"' ")
    for k in intents:
        if reply == k.lower() or reply.startswith(k.lower()):
            return k
    return default


class UniversalRetrievalConversation:
    def __init__(self, llm: Callable[[str], str], make_index: Callable[[], object] = BM25Index, k: int = 3,
                 min_score: float = 0.05, max_turns: int = 4, rewrite: bool = False,
                 intents: Optional[Dict[str, str]] = None, intent_replies: Optional[Dict[str, str]] = None) -> None:
        self.llm = llm
        self.intents, self.intent_replies = intents, intent_replies or {}
        self.chains: Dict[str, RetrievalQA] = {}
        for lang in ("en", "zh"):
            chain = RetrievalQA(make_index(), llm, k=k, min_score=min_score, memory=ConversationMemory(max_turns),
                                rewrite=rewrite)
            chain.PROMPT = PROMPTS[lang]
            self.chains[lang] = chain

    def add_documents(self, texts: Sequence[str], source: str = "default", chunk_size: int = 256) -> Dict[str, int]:
        added = {"en": 0, "zh": 0}
        for t in texts:
            lang = detect_language(t)
            chunks = split_chinese_text(t, chunk_size // 2) if lang == "zh" else split_text(t, chunk_size, chunk_size // 8)
            added[lang] += self.chains[lang].index.add_documents(chunks, source=source)
        return added

    def run(self, question: str) -> Tuple[str, List[Dict], Dict[str, str]]:
        """(answer, retrieved source chunks, {"language", "intent"})."""
        lang = detect_language(question)
        meta = {"language": lang, "intent": ""}
        if self.intents:
            meta["intent"] = classify_intent(self.llm, question, self.intents)
            if meta["intent"] in self.intent_replies:              # canned flow: no retrieval, still remembered
                answer = self.intent_replies[meta["intent"]]
                self.chains[lang].memory.add(question, answer)
                return answer, [], meta
        chain = self.chains[lang]
        prompt, sources = chain.build_prompt(question)
        answer = chain.generate(prompt) if sources else REFUSAL[lang]
        chain.memory.add(question, answer)
        return answer, sources, meta

    def reset(self) -> None:
        for c in self.chains.values():
            c.memory = ConversationMemory(c.memory.max_turns)
