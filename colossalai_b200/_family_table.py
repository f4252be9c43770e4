"""Entry points of the zoo's text families, as pure data (no imports from the package: both the model zoo -
`models/families.py` - and the policy zoo - `shardformer/policies/zoo.py` - are generated from it).

One row per family: the preset used when a class is built without a config, a one-line description, and the HF-style
class names with the head each one carries (`backbone` = hidden states, `lm` = causal LM head, `mlm` = masked LM,
`seq_cls` / `tok_cls` / `qa` / `choice` = the classification heads of `models/heads.py`).
Parity: the class lists of the reference's `shardformer/policies/auto_policy.py:_POLICY_LIST`."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple

__all__ = ["Family", "FAMILIES", "EXTRA_POLICY_FAMILIES", "family_classes"]


@dataclass(frozen=True)
class Family:
    preset: str                               # config used when a class is built without one
    summary: str
    classes: Tuple[Tuple[str, str], ...]      # (HF-style class name, head kind)
    defaults: Dict[str, object] = field(default_factory=dict)      # config fields the family always sets


def _decoder(prefix: str, *extra: Tuple[str, str]) -> Tuple[Tuple[str, str], ...]:
    return ((f"{prefix}Model", "backbone"), (f"{prefix}ForCausalLM", "lm")) + tuple(extra)


FAMILIES: Dict[str, Family] = {
    "llama": Family("llama2-7b", "Llama / Llama-2 / Llama-3 (RMSNorm, RoPE, SwiGLU, GQA)",
                    _decoder("Llama", ("LlamaForSequenceClassification", "seq_cls"))),
    "mistral": Family("mistral-7b", "Mistral (Llama block + sliding-window attention)",
                      _decoder("Mistral", ("MistralForSequenceClassification", "seq_cls"))),
    "qwen2": Family("qwen2-7b", "Qwen2 (Llama block with QKV bias)",
                    _decoder("Qwen2", ("Qwen2ForSequenceClassification", "seq_cls")),
                    {"attention_bias": True, "attention_out_bias": False}),
    "qwen3": Family("qwen3-8b", "Qwen3 (per-head q/k RMSNorm, no QKV bias)",
                    _decoder("Qwen3", ("Qwen3ForSequenceClassification", "seq_cls")), {"qk_norm": True}),
    "mixtral": Family("mixtral-8x7b", "Mixtral sparse MoE (top-2 of 8 experts, expert parallel)", _decoder("Mixtral")),
    "deepseek": Family("deepseek-moe-16b",
                       "DeepSeekMoE (fine-grained routed experts + shared experts, leading dense layers)",
                       _decoder("Deepseek")),
    "deepseek_v3": Family("deepseek-tiny",
                          "DeepSeek-V3 routing (sigmoid scores, group-limited top-k, routed scaling) + MLA",
                          _decoder("DeepseekV3")),
    "chatglm": Family("chatglm2-6b", "ChatGLM2/3 (RMSNorm, SwiGLU, multi-query groups, interleaved half-rotary RoPE)",
                      (("ChatGLMModel", "backbone"), ("ChatGLMForConditionalGeneration", "lm"))),
    "command": Family("command-r", "Cohere Command-R (parallel block, bias-free LayerNorm, logit scale, tied embeddings)",
                      _decoder("Cohere")),
    "gpt2": Family("gpt2", "GPT-2 (learned positions, LayerNorm, GELU MLP, tied embeddings)",
                   (("GPT2Model", "backbone"), ("GPT2LMHeadModel", "lm"), ("GPT2DoubleHeadsModel", "lm"),
                    ("GPT2ForQuestionAnswering", "qa"), ("GPT2ForTokenClassification", "tok_cls"),
                    ("GPT2ForSequenceClassification", "seq_cls"))),
    "gptj": Family("gptj-6b", "GPT-J (parallel attention + MLP block, interleaved partial RoPE)",
                   _decoder("GPTJ", ("GPTJForSequenceClassification", "seq_cls"), ("GPTJForQuestionAnswering", "qa"))),
    "opt": Family("opt-125m", "OPT (learned positions, LayerNorm, ReLU MLP)",
                  _decoder("OPT", ("OPTForSequenceClassification", "seq_cls"), ("OPTForQuestionAnswering", "qa"))),
    "bloom": Family("bloom-560m", "BLOOM (ALiBi, LayerNorm, embedding LayerNorm)",
                    _decoder("Bloom", ("BloomForSequenceClassification", "seq_cls"),
                             ("BloomForTokenClassification", "tok_cls"), ("BloomForQuestionAnswering", "qa"))),
    "falcon": Family("falcon-7b", "Falcon (parallel block, multi-query / grouped attention)",
                     _decoder("Falcon", ("FalconForSequenceClassification", "seq_cls"),
                              ("FalconForTokenClassification", "tok_cls"), ("FalconForQuestionAnswering", "qa"))),
    "bert": Family("bert-base", "BERT encoder (post-LN, learned positions + token types, bidirectional)",
                   (("BertModel", "backbone"), ("BertForPreTraining", "mlm"), ("BertLMHeadModel", "mlm"),
                    ("BertForMaskedLM", "mlm"), ("BertForSequenceClassification", "seq_cls"),
                    ("BertForTokenClassification", "tok_cls"), ("BertForNextSentencePrediction", "seq_cls"),
                    ("BertForMultipleChoice", "choice"), ("BertForQuestionAnswering", "qa"))),
}


def family_classes(family: str) -> Tuple[str, ...]:
    """Entry-point class names of a family, in table order."""
    return tuple(name for name, _ in FAMILIES[family].classes)


# families that are ordinary modules (behaviour of their own) but whose zoo policies are the generic one
EXTRA_POLICY_FAMILIES: Dict[str, Tuple[str, ...]] = {
    "baichuan": ("BaichuanModel", "BaichuanForCausalLM", "BaichuanForSequenceClassification"),
}
