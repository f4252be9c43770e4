"""Retrieval-augmented QA (reference `applications/ColossalQA/colossalqa`): text splitter, embedding index with
cosine retrieval, conversation memory and the prompt assembly around any generator callable."""
from .rag import ConversationMemory, EmbeddingIndex, RetrievalQA, hashing_embedder, split_text

__all__ = ["split_text", "EmbeddingIndex", "ConversationMemory", "RetrievalQA", "hashing_embedder"]
