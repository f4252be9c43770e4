"""Continual pre-training / SFT utilities (reference `applications/Colossal-LLaMA/colossal_llama`): spliced constant-
length dataset, vocabulary expansion, resumable checkpoints, streaming chat."""
from .dataset import ClosedToConstantLengthSplicedDataset, supervised_tokenize_pretrain
from .tuning import (activate_neftune, deactivate_neftune, format_numel_str, freeze_non_embeds_parameters,
                     get_model_numel, plan_vocab_expansion, unfreeze_parameters)
from .utils import expand_vocab, load_checkpoint, save_checkpoint, stream_chat

__all__ = ["ClosedToConstantLengthSplicedDataset", "supervised_tokenize_pretrain", "expand_vocab", "save_checkpoint",
           "load_checkpoint", "stream_chat", "activate_neftune", "deactivate_neftune", "freeze_non_embeds_parameters",
           "unfreeze_parameters", "plan_vocab_expansion", "get_model_numel", "format_numel_str"]
