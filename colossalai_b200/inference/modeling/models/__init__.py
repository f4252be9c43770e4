from .glide_llama import GlideLlamaConfig, GlideLlamaForCausalLM
from .nopadding import NopadBaichuanForCausalLM, NopadLlamaForCausalLM, build_nopad_model

__all__ = ["GlideLlamaConfig", "GlideLlamaForCausalLM", "NopadLlamaForCausalLM", "NopadBaichuanForCausalLM",
           "build_nopad_model"]
