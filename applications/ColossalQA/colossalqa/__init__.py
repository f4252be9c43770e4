"""Retrieval-augmented QA (reference `applications/ColossalQA/colossalqa`): text splitter, embedding index with
cosine retrieval, conversation memory and the prompt assembly around any generator callable."""
from .conversation import (LocalLLM, UniversalRetrievalConversation, classify_intent, detect_language, load_table,
                           split_chinese_text)
from .rag import ConversationMemory, EmbeddingIndex, RetrievalQA, hashing_embedder, split_text
from .retrieval import BM25Index, HybridRetriever, load_documents, rewrite_follow_up, tokenize

__all__ = ["split_text", "EmbeddingIndex", "ConversationMemory", "RetrievalQA", "hashing_embedder", "BM25Index",
           "HybridRetriever", "load_documents", "rewrite_follow_up", "tokenize", "LocalLLM", "UniversalRetrievalConversation",
           "classify_intent", "detect_language", "load_table", "split_chinese_text"]
