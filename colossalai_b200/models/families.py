"""The text families of the zoo, as data.

Every decoder / encoder family runs on ONE parallel-aware backbone (`transformer.py`, heads in `heads.py`); what makes a
family is its `ModelConfig` switches (`config.py`).  A family therefore needs no module of hand-written subclasses:
this table names its HF-style entry points, the head each one carries and the preset used when it is built without a
config, and `_install()` materialises them as real classes in real (synthetic) modules, so

    from colossalai_b200.models.llama import LlamaForCausalLM          # works, as does pickling / `__qualname__`
    LlamaForCausalLM()                                                 # llama2-7b shape
    LlamaForCausalLM(get_config("llama3-8b", num_hidden_layers=2))

and the auto-policy lookup by qualified class name (`shardformer/policies/auto_policy.py`) keep working.  The table
itself is `colossalai_b200/_family_table.py` (pure data, shared with the policy zoo).  Families with
behaviour of their own (Baichuan's `NormHead`, ViT, T5, Whisper, BLIP-2, SAM, DiT) stay ordinary modules.

Parity: the reference's per-family `shardformer/modeling/<family>.py` forwards + `policies/<family>.py` class lists
(llama.py:30-400, mistral.py, qwen2.py, qwen3.py, mixtral.py:54-208, deepseek.py:63-230, deepseek_v3.py:26, chatglm2.py,
command.py, gpt2.py, gptj.py, opt.py, bloom.py, falcon.py, bert.py).
"""
from __future__ import annotations

import sys
import types

from .._family_table import FAMILIES, Family, family_classes
from .config import ModelConfig, get_config
from .heads import (TransformerBackboneModel, TransformerForMaskedLM, TransformerForMultipleChoice,
                    TransformerForQuestionAnswering, TransformerForSequenceClassification,
                    TransformerForTokenClassification)
from .transformer import TransformerLMHeadModel

HEADS = {
    "backbone": TransformerBackboneModel,
    "lm": TransformerLMHeadModel,
    "mlm": TransformerForMaskedLM,
    "seq_cls": TransformerForSequenceClassification,
    "tok_cls": TransformerForTokenClassification,
    "qa": TransformerForQuestionAnswering,
    "choice": TransformerForMultipleChoice,
}


def _make_class(module_name: str, fam: Family, name: str, kind: str) -> type:
    base = HEADS[kind]

    def __init__(self, config: ModelConfig = None, **kw) -> None:
        base.__init__(self, config if config is not None else get_config(fam.preset), **kw)

    return type(name, (base,), {
        "__init__": __init__, "__module__": module_name, "__qualname__": name,
        "__doc__": f"{name}: `{base.__name__}` for the family \"{fam.summary}\"; built without a config it takes the "
                   f"`{fam.preset}` preset."})


def _install() -> None:
    """Create one module per family (`colossalai_b200.models.<family>`) holding its classes."""
    pkg = __name__.rsplit(".", 1)[0]
    parent = sys.modules[pkg]
    for key, fam in FAMILIES.items():
        name = f"{pkg}.{key}"
        mod = types.ModuleType(name, f"{fam.summary}.  Generated from `models/families.py`.")
        mod.DEFAULT_PRESET = fam.preset
        mod.FAMILY_DEFAULTS = dict(fam.defaults)
        mod.default_config = (lambda preset: lambda **overrides: get_config(preset, **overrides))(fam.preset)
        mod.default_config.__doc__ = "The family's reference-size config (override any field, e.g. `num_hidden_layers=2`)."
        exported = ["default_config"]
        for cls_name, kind in fam.classes:
            setattr(mod, cls_name, _make_class(name, fam, cls_name, kind))
            exported.append(cls_name)
        mod.__all__ = exported
        mod.__package__ = pkg
        sys.modules[name] = mod
        setattr(parent, key, mod)


_install()
