"""Model zoo.  One parallel-aware backbone (`transformer.py`, heads in `heads.py`); the text families are rows of
`_family_table.py` turned into classes by `families.py`, the vision / encoder-decoder / multimodal ones are modules."""
from .config import MODEL_ZOO, ModelConfig, MoEConfig, get_config
from .heads import (
    TransformerBackboneModel,
    TransformerForMaskedLM,
    TransformerForMultipleChoice,
    TransformerForQuestionAnswering,
    TransformerForSequenceClassification,
    TransformerForTokenClassification,
)
from .transformer import Attention, DecoderLayer, MLP, SeqMeta, TransformerLMHeadModel, TransformerModel
from . import families  # noqa: F401  (generates `models.<family>` for the table-driven families)
from .families import FAMILIES
from .llama import LlamaForCausalLM, LlamaForSequenceClassification, LlamaModel
from .gpt2 import GPT2LMHeadModel, GPT2Model
from .mixtral import MixtralForCausalLM
from .deepseek import DeepseekForCausalLM


_EXTRA_FAMILIES = [("vit", "VIT_ZOO", "ViTForImageClassification"), ("t5", "T5_ZOO", "T5ForConditionalGeneration"),
                   ("whisper", "WHISPER_ZOO", "WhisperForConditionalGeneration"),
                   ("blip2", "BLIP2_ZOO", "Blip2ForConditionalGeneration"), ("sam", "SAM_ZOO", "SamModel")]


def build_model(name_or_config, **overrides):
    """`build_model("llama3-8b")` or `build_model(ModelConfig(...))` -> causal LM of the right family class."""
    import importlib

    # vision / encoder-decoder / multimodal families carry their own config types and zoos
    for mod, zoo, cls in _EXTRA_FAMILIES:
        m = None
        if isinstance(name_or_config, str):
            m = importlib.import_module(f"colossalai_b200.models.{mod}")
            z = getattr(m, zoo)
            if name_or_config in z:
                cfg = z[name_or_config]
                return getattr(m, cls)(cfg.replace(**overrides) if overrides else cfg)
        elif getattr(name_or_config, "model_type", None) == mod:
            m = importlib.import_module(f"colossalai_b200.models.{mod}")
            return getattr(m, cls)(name_or_config)
    cfg = get_config(name_or_config, **overrides) if isinstance(name_or_config, str) else name_or_config

    fam = {"chatglm": ("chatglm", "ChatGLMForConditionalGeneration"), "gpt2": ("gpt2", "GPT2LMHeadModel"),
           "bert": ("bert", "BertForMaskedLM"), "command": ("command", "CohereForCausalLM"),
           "deepseek": ("deepseek", "DeepseekForCausalLM"), "deepseek_v3": ("deepseek_v3", "DeepseekV3ForCausalLM"),
           "gptj": ("gptj", "GPTJForCausalLM"), "opt": ("opt", "OPTForCausalLM")}
    mod, cls = fam.get(cfg.model_type, (cfg.model_type, cfg.model_type.capitalize() + "ForCausalLM"))
    try:
        return getattr(importlib.import_module(f"colossalai_b200.models.{mod}"), cls)(cfg)
    except (ImportError, AttributeError):
        return TransformerLMHeadModel(cfg)


__all__ = ["MODEL_ZOO", "ModelConfig", "MoEConfig", "get_config", "build_model", "TransformerBackboneModel",
           "TransformerForMaskedLM", "TransformerForMultipleChoice", "TransformerForQuestionAnswering",
           "TransformerForSequenceClassification", "TransformerForTokenClassification", "Attention", "DecoderLayer",
           "MLP", "SeqMeta", "TransformerLMHeadModel", "TransformerModel", "LlamaForCausalLM",
           "LlamaForSequenceClassification", "LlamaModel", "GPT2LMHeadModel", "GPT2Model", "MixtralForCausalLM",
           "DeepseekForCausalLM"]
