"""Policy registry + auto lookup.  Parity: reference `colossalai/shardformer/policies/auto_policy.py`
(`_POLICY_LIST`, `get_autopolicy`, `import_policy`)."""
from __future__ import annotations

import importlib
from dataclasses import dataclass
from typing import Dict

import torch.nn as nn

from .base_policy import Policy

__all__ = ["PolicyLocation", "get_autopolicy", "import_policy", "register_policy", "_POLICY_LIST"]


@dataclass
class PolicyLocation:
    file_name: str
    class_name: str


_P = "colossalai_b200.shardformer.policies"

# fully-qualified model class name -> policy location.  All families of our zoo share one backbone, hence one
# policy implementation with family-specific subclasses kept for discoverability / custom overrides.
_POLICY_LIST: Dict[str, PolicyLocation] = {}


def register_policy(qualname: str, file_name: str, class_name: str) -> None:
    _POLICY_LIST[qualname] = PolicyLocation(file_name, class_name)


for _cls in ("TransformerLMHeadModel", "TransformerModel"):
    register_policy(f"colossalai_b200.models.transformer.{_cls}", "transformer", "TransformerForCausalLMPolicy")

# text families of the zoo: classes and policies are both generated from `_family_table.py` (`models/families.py`,
# `policies/zoo.py`); the policy of model class `X` of family `f` is `policies.<f>.XPolicy`
from ..._family_table import EXTRA_POLICY_FAMILIES, FAMILIES, family_classes  # noqa: E402

for _fam in FAMILIES:
    for _c in family_classes(_fam):
        register_policy(f"colossalai_b200.models.{_fam}.{_c}", _fam, f"{_c}Policy")
for _fam, _classes in EXTRA_POLICY_FAMILIES.items():
    for _c in _classes:
        register_policy(f"colossalai_b200.models.{_fam}.{_c}", _fam, f"{_c}Policy")

# families with policies of their own (vision / encoder-decoder / multimodal)
_OWN_POLICY_FAMILIES = {
    "vit": ["ViTModel", "ViTForImageClassification", "ViTForMaskedImageModeling"],
    "t5": ["T5Model", "T5ForConditionalGeneration", "T5EncoderModel", "T5ForTokenClassification"],
    "whisper": ["WhisperModel", "WhisperForConditionalGeneration", "WhisperForAudioClassification"],
    "blip2": ["Blip2Model", "Blip2ForConditionalGeneration"],
    "sam": ["SamModel"],
}
for _fam, _classes in _OWN_POLICY_FAMILIES.items():
    for _c in _classes:
        register_policy(f"colossalai_b200.models.{_fam}.{_c}", _fam, f"{_c}Policy")


# user-supplied HuggingFace modules (sharded in place by sub-module / method replacement)
for _mod, _pre in (("llama", "Llama"), ("mistral", "Mistral"), ("qwen2", "Qwen2"), ("qwen3", "Qwen3"), ("cohere", "Cohere"),
                   ("glm", "Glm")):
    for _suffix in ("Model", "ForCausalLM"):
        register_policy(f"transformers.models.{_mod}.modeling_{_mod}.{_pre}{_suffix}", "hf_decoder", "HFDecoderPolicy")
for _c in ("GPT2Model", "GPT2LMHeadModel"):
    register_policy(f"transformers.models.gpt2.modeling_gpt2.{_c}", "hf_gpt", "HFGPT2Policy")
for _c in ("OPTModel", "OPTForCausalLM"):
    register_policy(f"transformers.models.opt.modeling_opt.{_c}", "hf_gpt", "HFOPTPolicy")
for _mod, _pre in (("mixtral", "Mixtral"), ("qwen3_moe", "Qwen3Moe"), ("qwen2_moe", "Qwen2Moe"),
                   ("deepseek_v3", "DeepseekV3"), ("deepseek_v2", "DeepseekV2")):
    for _suffix in ("Model", "ForCausalLM"):
        register_policy(f"transformers.models.{_mod}.modeling_{_mod}.{_pre}{_suffix}", "hf_moe", f"HF{_pre}Policy")
for _c in ("FalconModel", "FalconForCausalLM"):
    register_policy(f"transformers.models.falcon.modeling_falcon.{_c}", "hf_gpt", "HFFalconPolicy")
for _c in ("BloomModel", "BloomForCausalLM"):
    register_policy(f"transformers.models.bloom.modeling_bloom.{_c}", "hf_gpt", "HFBloomPolicy")
for _c in ("GPTJModel", "GPTJForCausalLM"):
    register_policy(f"transformers.models.gptj.modeling_gptj.{_c}", "hf_gpt", "HFGPTJPolicy")
for _c in ("BertModel", "BertForSequenceClassification", "BertForTokenClassification", "BertForQuestionAnswering",
           "BertForMultipleChoice", "BertForNextSentencePrediction"):
    register_policy(f"transformers.models.bert.modeling_bert.{_c}", "hf_encoder", "HFBertPolicy")
for _c in ("WhisperModel", "WhisperForConditionalGeneration"):
    register_policy(f"transformers.models.whisper.modeling_whisper.{_c}", "hf_encdec", "HFWhisperPolicy")
for _c in ("T5Model", "T5ForConditionalGeneration", "T5EncoderModel"):
    register_policy(f"transformers.models.t5.modeling_t5.{_c}", "hf_encdec", "HFT5Policy")
for _c in ("ViTModel", "ViTForImageClassification"):
    register_policy(f"transformers.models.vit.modeling_vit.{_c}", "hf_encoder", "HFViTPolicy")


for _c in ("SamModel", "SamVisionModel"):
    register_policy(f"transformers.models.sam.modeling_sam.{_c}", "hf_vision", "HFSamPolicy")
for _c in ("Blip2Model", "Blip2ForConditionalGeneration"):
    register_policy(f"transformers.models.blip_2.modeling_blip_2.{_c}", "hf_vision", "HFBlip2Policy")


def import_policy(loc: PolicyLocation) -> type:
    from . import zoo  # noqa: F401  (installs the generated per-family policy modules)

    module = importlib.import_module(f"{_P}.{loc.file_name}")
    return getattr(module, loc.class_name)


def _fullname(obj) -> str:
    klass = obj.__class__
    module = klass.__module__
    return klass.__qualname__ if module == "builtins" else module + "." + klass.__qualname__


def get_autopolicy(model: nn.Module) -> Policy:
    name = _fullname(model)
    loc = _POLICY_LIST.get(name)
    if loc is None:
        # subclasses of our generic models inherit the generic policy
        for klass in model.__class__.__mro__:
            key = klass.__module__ + "." + klass.__qualname__
            if key in _POLICY_LIST:
                loc = _POLICY_LIST[key]
                break
    if loc is None:
        raise NotImplementedError(
            f"auto policy for {name} is not implemented; supported: {sorted(_POLICY_LIST.keys())}")
    return import_policy(loc)()
