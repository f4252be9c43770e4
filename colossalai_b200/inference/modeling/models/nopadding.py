"""No-padding inference models.  In the reference these are separate forwards patched over HF modules
(`inference/modeling/models/nopadding_llama.py:35-677`, `nopadding_baichuan.py:1-420`: fused qkv weight, RMSNorm with
fused residual add, un-padded 1-D token stream + `InputMetaData`).  Our generic transformer is already token-major,
keeps one fused QKV GEMM and one fused gate|up GEMM per block and takes the paged-KV runtime as `kv_cache`, so the
no-padding classes only pin the family configuration and expose the engine-facing call."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ....models.baichuan import BaichuanForCausalLM
from ....models.config import ModelConfig
from ....models.llama import LlamaForCausalLM
from ....models.transformer import SeqMeta

__all__ = ["NopadLlamaForCausalLM", "NopadBaichuanForCausalLM", "build_nopad_model"]


class _NopadMixin:
    @torch.inference_mode()
    def infer(self, input_tokens_ids: torch.Tensor, positions: torch.Tensor, kv_runtime) -> torch.Tensor:
        """`input_tokens_ids` [tokens] un-padded; `kv_runtime` already `set_step()`-ed.  Returns logits [tokens, V]."""
        meta = SeqMeta(batch=1, seqlen=input_tokens_ids.numel(), positions=positions)
        out: Dict[str, torch.Tensor] = self(input_ids=input_tokens_ids.view(1, -1), kv_cache=kv_runtime, meta=meta)
        return out["logits"]


class NopadLlamaForCausalLM(_NopadMixin, LlamaForCausalLM):
    pass


class NopadBaichuanForCausalLM(_NopadMixin, BaichuanForCausalLM):
    pass


def build_nopad_model(cfg: ModelConfig):
    return (NopadBaichuanForCausalLM if cfg.model_type == "baichuan" else NopadLlamaForCausalLM)(cfg)
