"""Every decoder / encoder family of the zoo under TP=2 (and TP+SP) against its unsharded self: loss and gathered
gradients — exercises the family switches inside the sharded layers (ALiBi slopes per TP rank, parallel blocks,
per-head q/k norms, partial / interleaved rotary, LM-head bias, post-norm encoders, tied embeddings).
Reference: tests/test_shardformer/test_model/test_shard_{bloom,opt,gptj,falcon,qwen2,command,chatglm2,bert,...}.py."""
import pytest
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn

FAMILIES = ["mistral-tiny", "qwen2-tiny", "qwen3-tiny", "opt-tiny", "bloom-tiny", "falcon-tiny", "gptj-tiny",
            "chatglm-tiny", "command-tiny", "bert-tiny", "baichuan-tiny"]


def _worker(rank, world_size, port):
    from test_shard_llama import _run_one

    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    _run_one("deepseek_v3-tiny", dict(tp=2, sp_mode=None), atol=5e-5)     # MLA: head-split up-projections, shared rope key
    for name in FAMILIES:
        _run_one(name, dict(tp=2, sp_mode=None), atol=5e-5)
        if name not in ("bloom-tiny", "baichuan-tiny", "bert-tiny"):      # (ALiBi / padded-mask paths build full masks)
            _run_one(name, dict(tp=2, sp_mode="split_gather"), atol=5e-5)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_all_families_tp2():
    import os
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    spawn(_worker, 2)
