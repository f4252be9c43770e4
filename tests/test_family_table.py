"""The table-driven family zoo: every row yields importable, picklable classes whose auto-policy resolves, and the
generated modules behave like the hand-written ones they replace (default preset, `default_config`, `__all__`)."""
import importlib
import pickle

import pytest

from colossalai_b200._family_table import EXTRA_POLICY_FAMILIES, FAMILIES, family_classes
from colossalai_b200.models import get_config
from colossalai_b200.models.families import HEADS
from colossalai_b200.shardformer.policies.auto_policy import _POLICY_LIST, import_policy
from colossalai_b200.shardformer.policies.transformer import TransformerPolicy


@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_family_modules_are_generated(family):
    mod = importlib.import_module(f"colossalai_b200.models.{family}")
    pol = importlib.import_module(f"colossalai_b200.shardformer.policies.{family}")
    fam = FAMILIES[family]
    assert mod.DEFAULT_PRESET == fam.preset and set(family_classes(family)) | {"default_config"} == set(mod.__all__)
    assert mod.default_config(num_hidden_layers=1).num_hidden_layers == 1
    for name, kind in fam.classes:
        cls = getattr(mod, name)
        assert issubclass(cls, HEADS[kind]) and cls.__module__ == mod.__name__
        assert pickle.loads(pickle.dumps(cls)) is cls                   # resolvable by qualified name
        loc = _POLICY_LIST[f"{mod.__name__}.{name}"]
        policy_cls = import_policy(loc)
        assert policy_cls is getattr(pol, f"{name}Policy") and issubclass(policy_cls, TransformerPolicy)


def test_generated_class_builds_with_and_without_config():
    from colossalai_b200.models.llama import LlamaForSequenceClassification
    from colossalai_b200.models.qwen3 import Qwen3ForCausalLM
    from colossalai_b200.shardformer.policies import get_autopolicy

    m = Qwen3ForCausalLM(get_config("qwen3-tiny"))
    assert m.cfg.qk_norm and type(get_autopolicy(m)).__name__ == "Qwen3ForCausalLMPolicy"
    cfg = get_config("llama-tiny")
    m = LlamaForSequenceClassification(cfg, num_labels=3)
    import torch

    out = m(input_ids=torch.randint(0, cfg.vocab_size, (2, 8)))
    assert out["logits"].shape == (2, 3)


def test_custom_policy_subclass_of_generated_policy():
    """The reason the per-class policy names exist: a user subclasses exactly one and passes it as `custom_policy`."""
    from colossalai_b200.shardformer.policies.gpt2 import GPT2LMHeadModelPolicy

    class MyPolicy(GPT2LMHeadModelPolicy):
        def config_sanity_check(self):
            self.checked = True
            super().config_sanity_check()

    assert MyPolicy.__mro__[1] is GPT2LMHeadModelPolicy
    for fam, classes in EXTRA_POLICY_FAMILIES.items():
        pol = importlib.import_module(f"colossalai_b200.shardformer.policies.{fam}")
        assert all(hasattr(pol, f"{c}Policy") for c in classes)
