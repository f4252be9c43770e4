"""Continual pre-training / SFT utilities (reference `applications/Colossal-LLaMA/colossal_llama`): spliced constant-
length dataset, vocabulary expansion, resumable checkpoints, streaming chat."""
from .dataset import ClosedToConstantLengthSplicedDataset, supervised_tokenize_pretrain
from .utils import expand_vocab, load_checkpoint, save_checkpoint, stream_chat

__all__ = ["ClosedToConstantLengthSplicedDataset", "supervised_tokenize_pretrain", "expand_vocab", "save_checkpoint",
           "load_checkpoint", "stream_chat"]
