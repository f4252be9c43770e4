"""Hugging Face checkpoint <-> colossalai_b200 model conversion.

The reference shards HF `transformers` modules in place, so HF checkpoints load as-is (`checkpoint_io/*`,
`lazy/pretrained.py`).  Our generic transformer fuses q/k/v into `qkv_proj`, gate/up into `gate_up_proj` and batches the
experts of an MoE block, so this module maps names and fuses / splits tensors:

    model = load_hf_checkpoint("/path/to/hf_dir", dtype=torch.bfloat16)       # config.json + *.safetensors / *.bin
    sd    = to_hf_state_dict(model)                                            # export back to HF naming

Supported: llama / mistral / qwen2 / qwen3 / baichuan (llama-like), mixtral, gpt2, opt, bloom, falcon (7B-style
multi-query layout) and the bert backbone.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Iterator, Optional, Tuple

import torch

from .config import ModelConfig, MoEConfig

__all__ = ["config_from_hf", "convert_hf_state_dict", "load_hf_checkpoint", "to_hf_state_dict", "iter_hf_shards",
           "from_hf_model", "is_hf_model", "save_hf_checkpoint", "hf_config_dict"]

_LLAMA_LIKE = ("llama", "mistral", "qwen2", "qwen3", "mixtral", "baichuan", "command")


def config_from_hf(hf: dict) -> ModelConfig:
    mt = hf.get("model_type", "llama")
    if mt == "gpt2":
        return ModelConfig(model_type="gpt2", vocab_size=hf["vocab_size"], hidden_size=hf["n_embd"],
                           intermediate_size=hf.get("n_inner") or 4 * hf["n_embd"], num_hidden_layers=hf["n_layer"],
                           num_attention_heads=hf["n_head"], max_position_embeddings=hf["n_positions"],
                           pos_type="learned", norm_type="layer", hidden_act="gelu_new", glu=False,
                           attention_bias=True, mlp_bias=True, tie_word_embeddings=True,
                           norm_eps=hf.get("layer_norm_epsilon", 1e-5))
    if mt == "opt":
        assert hf.get("do_layer_norm_before", True) and hf.get("word_embed_proj_dim", hf["hidden_size"]) == hf["hidden_size"], \
            "OPT-350m style post-LN / projected embeddings are not supported"
        return ModelConfig(model_type="opt", vocab_size=hf["vocab_size"], hidden_size=hf["hidden_size"],
                           intermediate_size=hf["ffn_dim"], num_hidden_layers=hf["num_hidden_layers"],
                           num_attention_heads=hf["num_attention_heads"],
                           max_position_embeddings=hf["max_position_embeddings"], norm_type="layer",
                           hidden_act=hf.get("activation_function", "relu"), glu=False, attention_bias=True, mlp_bias=True,
                           pos_type="learned", tie_word_embeddings=hf.get("tie_word_embeddings", True), norm_eps=1e-5,
                           pad_token_id=hf.get("pad_token_id"), bos_token_id=hf.get("bos_token_id", 2),
                           eos_token_id=hf.get("eos_token_id", 2))
    if mt == "bloom":
        h = hf.get("hidden_size", hf.get("n_embed"))
        return ModelConfig(model_type="bloom", vocab_size=hf["vocab_size"], hidden_size=h, intermediate_size=4 * h,
                           num_hidden_layers=hf.get("n_layer", hf.get("num_hidden_layers")),
                           num_attention_heads=hf.get("n_head", hf.get("num_attention_heads")), norm_type="layer",
                           hidden_act="gelu_new", glu=False, attention_bias=True, mlp_bias=True, pos_type="alibi",
                           embed_norm=True, tie_word_embeddings=True, norm_eps=hf.get("layer_norm_epsilon", 1e-5))
    if mt == "falcon":
        assert not hf.get("new_decoder_architecture", False) and hf.get("parallel_attn", True) and not hf.get("alibi", False), \
            "only the falcon-7b style layout (parallel attention, rotary, single layer norm) is mapped"
        nh = hf["num_attention_heads"]
        nkv = 1 if hf.get("multi_query", True) else nh
        return ModelConfig(model_type="falcon", vocab_size=hf["vocab_size"], hidden_size=hf["hidden_size"],
                           intermediate_size=hf.get("ffn_hidden_size") or 4 * hf["hidden_size"],
                           num_hidden_layers=hf["num_hidden_layers"], num_attention_heads=nh, num_key_value_heads=nkv,
                           norm_type="layer", hidden_act="gelu", glu=False, parallel_block=True,
                           attention_bias=hf.get("bias", False), mlp_bias=hf.get("bias", False),
                           rope_theta=float(hf.get("rope_theta", 10000.0)),
                           max_position_embeddings=hf.get("max_position_embeddings", 2048),
                           tie_word_embeddings=hf.get("tie_word_embeddings", True),
                           norm_eps=hf.get("layer_norm_epsilon", 1e-5))
    if mt == "gptj":
        return ModelConfig(model_type="gptj", vocab_size=hf["vocab_size"], hidden_size=hf["n_embd"],
                           intermediate_size=hf.get("n_inner") or 4 * hf["n_embd"], num_hidden_layers=hf["n_layer"],
                           num_attention_heads=hf["n_head"], max_position_embeddings=hf.get("n_positions", 2048),
                           norm_type="layer", hidden_act=hf.get("activation_function", "gelu_new"), glu=False,
                           mlp_bias=True, parallel_block=True, rope_interleaved=True,
                           partial_rotary_factor=hf["rotary_dim"] / (hf["n_embd"] // hf["n_head"]), lm_head_bias=True,
                           norm_eps=hf.get("layer_norm_epsilon", 1e-5), tie_word_embeddings=False)
    if mt == "cohere":
        assert not hf.get("use_qk_norm", False), "Cohere checkpoints with per-head LayerNorm qk-norm are not mapped"
        return ModelConfig(model_type="command", vocab_size=hf["vocab_size"], hidden_size=hf["hidden_size"],
                           intermediate_size=hf["intermediate_size"], num_hidden_layers=hf["num_hidden_layers"],
                           num_attention_heads=hf["num_attention_heads"],
                           num_key_value_heads=hf.get("num_key_value_heads", hf["num_attention_heads"]),
                           max_position_embeddings=hf.get("max_position_embeddings", 8192), norm_type="layer",
                           parallel_block=True, tie_word_embeddings=hf.get("tie_word_embeddings", True),
                           logit_scale=hf.get("logit_scale", 0.0625), rope_theta=float(hf.get("rope_theta", 10000.0)),
                           rope_interleaved=True, norm_eps=hf.get("layer_norm_eps", 1e-5),
                           attention_bias=hf.get("attention_bias", False))
    if mt == "bert":
        return ModelConfig(model_type="bert", vocab_size=hf["vocab_size"], hidden_size=hf["hidden_size"],
                           intermediate_size=hf["intermediate_size"], num_hidden_layers=hf["num_hidden_layers"],
                           num_attention_heads=hf["num_attention_heads"],
                           max_position_embeddings=hf["max_position_embeddings"], norm_type="layer",
                           hidden_act=hf.get("hidden_act", "gelu"), glu=False, attention_bias=True, mlp_bias=True,
                           pos_type="learned", causal=False, type_vocab_size=hf.get("type_vocab_size", 2), post_norm=True,
                           final_norm=False, embed_norm=True, norm_eps=hf.get("layer_norm_eps", 1e-12))
    if mt == "deepseek_v3":
        rp = hf.get("rope_parameters") or {}
        scaling = hf.get("rope_scaling") or ({k: v for k, v in rp.items() if k != "rope_theta"}
                                              if rp.get("rope_type", "default") != "default" else None)
        return ModelConfig(
            model_type="deepseek_v3", vocab_size=hf["vocab_size"], hidden_size=hf["hidden_size"],
            intermediate_size=hf["intermediate_size"], num_hidden_layers=hf["num_hidden_layers"],
            num_attention_heads=hf["num_attention_heads"], max_position_embeddings=hf.get("max_position_embeddings", 4096),
            norm_eps=hf.get("rms_norm_eps", 1e-6), rope_theta=float(hf.get("rope_theta", rp.get("rope_theta", 10000.0))),
            rope_scaling=scaling, rope_interleaved=hf.get("rope_interleave", True),
            attention_bias=hf.get("attention_bias", False), tie_word_embeddings=hf.get("tie_word_embeddings", False),
            q_lora_rank=hf.get("q_lora_rank"), kv_lora_rank=hf["kv_lora_rank"], qk_nope_head_dim=hf["qk_nope_head_dim"],
            qk_rope_head_dim=hf["qk_rope_head_dim"], v_head_dim=hf["v_head_dim"],
            bos_token_id=hf.get("bos_token_id", 0), eos_token_id=hf.get("eos_token_id", 1),
            moe=MoEConfig(num_experts=hf["n_routed_experts"], top_k=hf["num_experts_per_tok"],
                          moe_intermediate_size=hf["moe_intermediate_size"],
                          n_shared_experts=hf.get("n_shared_experts") or 0,
                          first_k_dense_replace=hf.get("first_k_dense_replace", 0),
                          moe_layer_freq=hf.get("moe_layer_freq", 1), norm_topk_prob=hf.get("norm_topk_prob", True),
                          scoring_func="sigmoid", n_group=hf.get("n_group", 1), topk_group=hf.get("topk_group", 1),
                          routed_scaling_factor=hf.get("routed_scaling_factor", 1.0)))
    assert mt in _LLAMA_LIKE, f"unsupported HF model_type {mt!r}"
    kw = dict(model_type=mt, vocab_size=hf["vocab_size"], hidden_size=hf["hidden_size"],
              intermediate_size=hf["intermediate_size"], num_hidden_layers=hf["num_hidden_layers"],
              num_attention_heads=hf["num_attention_heads"],
              num_key_value_heads=hf.get("num_key_value_heads", hf["num_attention_heads"]),
              max_position_embeddings=hf.get("max_position_embeddings", 4096),
              rope_theta=float(hf.get("rope_theta", 10000.0)), norm_eps=hf.get("rms_norm_eps", 1e-6),
              tie_word_embeddings=hf.get("tie_word_embeddings", False), rope_scaling=hf.get("rope_scaling"))
    if hf.get("head_dim"):
        kw["head_dim"] = hf["head_dim"]
    if hf.get("sliding_window") and (mt in ("mistral", "mixtral") or hf.get("use_sliding_window", False)):
        kw["sliding_window"] = int(hf["sliding_window"])
    if mt == "qwen2":
        kw["attention_bias"], kw["attention_out_bias"] = True, False
    if mt == "qwen3":
        kw["qk_norm"] = True
    if mt == "baichuan":
        # 13B checkpoints (hidden 5120) use ALiBi instead of RoPE; Baichuan-2 (vocab 125696) has the NormHead
        kw["pos_type"] = "alibi" if hf["hidden_size"] >= 5120 else "rope"
        kw["norm_head"] = hf["vocab_size"] == 125696
        kw["max_position_embeddings"] = hf.get("max_position_embeddings", hf.get("model_max_length", 4096))
    if mt == "mixtral":
        kw["moe"] = MoEConfig(num_experts=hf["num_local_experts"], top_k=hf["num_experts_per_tok"])
    fields = ModelConfig.__dataclass_fields__
    return ModelConfig(**{k: v for k, v in kw.items() if k in fields})


def iter_hf_shards(path: str) -> Iterator[Dict[str, torch.Tensor]]:
    files = sorted(f for f in os.listdir(path) if f.endswith((".safetensors", ".bin")) and "optimizer" not in f
                   and not f.startswith("training_args"))
    for f in files:
        full = os.path.join(path, f)
        if f.endswith(".safetensors"):
            from safetensors.torch import load_file

            yield load_file(full)
        else:
            yield torch.load(full, map_location="cpu", weights_only=True)


def convert_hf_state_dict(hf_sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> Dict[str, torch.Tensor]:
    """HF names -> ours (fusing q/k/v, gate/up, stacking experts).  Keys it does not know are passed through."""
    out: Dict[str, torch.Tensor] = {}
    if cfg.model_type == "gpt2":
        for k, v in hf_sd.items():
            k = k[len("transformer."):] if k.startswith("transformer.") else k
            m = re.match(r"h\.(\d+)\.(.*)", k)
            if m:
                i, rest = m.group(1), m.group(2)
                pre = f"model.layers.{i}."
                table = {"ln_1.": "input_layernorm.", "ln_2.": "post_attention_layernorm.",
                         "attn.c_attn.": "self_attn.qkv_proj.", "attn.c_proj.": "self_attn.o_proj.",
                         "mlp.c_fc.": "mlp.up_proj.", "mlp.c_proj.": "mlp.down_proj."}
                for a, b in table.items():
                    if rest.startswith(a):
                        t = v.t().contiguous() if (rest.endswith("weight") and ".c_" in rest) else v   # Conv1D -> Linear
                        out[pre + b + rest[len(a):]] = t
                        break
            elif k == "wte.weight":
                out["model.embed_tokens.weight"] = v
            elif k == "wpe.weight":
                out["model.embed_positions.weight"] = v
            elif k.startswith("ln_f."):
                out["model.norm." + k[5:]] = v
            elif k == "lm_head.weight":
                out["lm_head.weight"] = v
        return out
    if cfg.model_type in _FAMILY_CONVERTERS:
        return _FAMILY_CONVERTERS[cfg.model_type](hf_sd, cfg)
    qkv: Dict[Tuple[str, str], Dict[str, torch.Tensor]] = {}
    gu: Dict[str, Dict[str, torch.Tensor]] = {}
    experts: Dict[str, Dict[int, Dict[str, torch.Tensor]]] = {}
    for k, v in hf_sd.items():
        m = re.match(r"(model\.layers\.\d+\.self_attn\.)([qkv])_proj\.(weight|bias)", k)
        if m:
            qkv.setdefault((m.group(1), m.group(3)), {})[m.group(2)] = v
            continue
        m = re.match(r"(model\.layers\.\d+\.self_attn\.)W_pack\.(weight|bias)", k)
        if m:       # Baichuan: already fused [q; k; v]
            out[f"{m.group(1)}qkv_proj.{m.group(2)}"] = v
            continue
        m = re.match(r"(model\.layers\.\d+\.mlp\.)(gate|up)_proj\.weight", k)
        if m:
            gu.setdefault(m.group(1), {})[m.group(2)] = v
            continue
        m = re.match(r"(model\.layers\.\d+\.)block_sparse_moe\.experts\.(\d+)\.(w[123])\.weight", k)
        if m:
            experts.setdefault(m.group(1), {}).setdefault(int(m.group(2)), {})[m.group(3)] = v
            continue
        m = re.match(r"(model\.layers\.\d+\.)block_sparse_moe\.gate\.weight", k)
        if m:
            out[m.group(1) + "mlp.router.gate.weight"] = v
            continue
        out[k] = v
    for (pre, kind), parts in qkv.items():
        out[f"{pre}qkv_proj.{kind}"] = torch.cat([parts["q"], parts["k"], parts["v"]], dim=0)
    for pre, parts in gu.items():
        out[pre + "gate_up_proj.weight"] = torch.cat([parts["gate"], parts["up"]], dim=0)
    for pre, ex in experts.items():
        ids = sorted(ex)
        out[pre + "mlp.experts.w_up"] = torch.stack([torch.cat([ex[i]["w1"], ex[i]["w3"]], 0) for i in ids])
        out[pre + "mlp.experts.w_down"] = torch.stack([ex[i]["w2"] for i in ids])
    return out


def _strip(k: str, *prefixes: str) -> str:
    for p in prefixes:
        if k.startswith(p):
            return k[len(p):]
    return k


def _convert_opt(hf_sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    parts: Dict[Tuple[str, str], Dict[str, torch.Tensor]] = {}
    for k, v in hf_sd.items():
        k = _strip(k, "model.decoder.", "decoder.")
        m = re.match(r"layers\.(\d+)\.(.*)", k)
        if m:
            pre, rest = f"model.layers.{m.group(1)}.", m.group(2)
            mm = re.match(r"self_attn\.([qkv])_proj\.(weight|bias)", rest)
            if mm:
                parts.setdefault((pre, mm.group(2)), {})[mm.group(1)] = v
                continue
            for a, b in {"self_attn.out_proj.": "self_attn.o_proj.", "self_attn_layer_norm.": "input_layernorm.",
                         "final_layer_norm.": "post_attention_layernorm.", "fc1.": "mlp.up_proj.",
                         "fc2.": "mlp.down_proj."}.items():
                if rest.startswith(a):
                    out[pre + b + rest[len(a):]] = v
                    break
        elif k == "embed_tokens.weight":
            out["model.embed_tokens.weight"] = v
        elif k == "embed_positions.weight":
            out["model.embed_positions.weight"] = v[2:]          # OPT reserves two offset rows
        elif k.startswith("final_layer_norm."):
            out["model.norm." + k[len("final_layer_norm."):]] = v
        elif k == "lm_head.weight":
            out["lm_head.weight"] = v
    for (pre, kind), p in parts.items():
        out[f"{pre}self_attn.qkv_proj.{kind}"] = torch.cat([p["q"], p["k"], p["v"]], 0)
    return out


def _convert_bloom(hf_sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    H, D = cfg.num_attention_heads, cfg.head_dim
    for k, v in hf_sd.items():
        k = _strip(k, "transformer.")
        m = re.match(r"h\.(\d+)\.(.*)", k)
        if m:
            pre, rest = f"model.layers.{m.group(1)}.", m.group(2)
            if rest.startswith("self_attention.query_key_value."):
                # BLOOM interleaves per head: rows are [head, (q, k, v), head_dim] -> ours is [q heads | k heads | v heads]
                kind = rest.rsplit(".", 1)[1]
                t = v.reshape(H, 3, D, *v.shape[1:]).transpose(0, 1).reshape(3 * H * D, *v.shape[1:])
                out[pre + "self_attn.qkv_proj." + kind] = t.contiguous()
                continue
            for a, b in {"self_attention.dense.": "self_attn.o_proj.", "mlp.dense_h_to_4h.": "mlp.up_proj.",
                         "mlp.dense_4h_to_h.": "mlp.down_proj.", "input_layernorm.": "input_layernorm.",
                         "post_attention_layernorm.": "post_attention_layernorm."}.items():
                if rest.startswith(a):
                    out[pre + b + rest[len(a):]] = v
                    break
        elif k == "word_embeddings.weight":
            out["model.embed_tokens.weight"] = v
        elif k.startswith("word_embeddings_layernorm."):
            out["model.embed_layernorm." + k.rsplit(".", 1)[1]] = v
        elif k.startswith("ln_f."):
            out["model.norm." + k[5:]] = v
        elif k == "lm_head.weight":
            out["lm_head.weight"] = v
    return out


def _convert_falcon(hf_sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for k, v in hf_sd.items():
        k = _strip(k, "transformer.")
        m = re.match(r"h\.(\d+)\.(.*)", k)
        if m:
            pre, rest = f"model.layers.{m.group(1)}.", m.group(2)
            if rest.startswith("self_attention.query_key_value.") and cfg.num_key_value_heads != 1:
                # multi-head falcon (rw-1b style) interleaves like BLOOM
                H, D = cfg.num_attention_heads, cfg.head_dim
                v = v.reshape(H, 3, D, *v.shape[1:]).transpose(0, 1).reshape(3 * H * D, *v.shape[1:]).contiguous()
            for a, b in {"self_attention.query_key_value.": "self_attn.qkv_proj.",
                         "self_attention.dense.": "self_attn.o_proj.", "mlp.dense_h_to_4h.": "mlp.up_proj.",
                         "mlp.dense_4h_to_h.": "mlp.down_proj.", "input_layernorm.": "input_layernorm."}.items():
                if rest.startswith(a):
                    out[pre + b + rest[len(a):]] = v
                    break
        elif k == "word_embeddings.weight":
            out["model.embed_tokens.weight"] = v
        elif k.startswith("ln_f."):
            out["model.norm." + k[5:]] = v
        elif k == "lm_head.weight":
            out["lm_head.weight"] = v
    return out


def _convert_bert(hf_sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    parts: Dict[Tuple[str, str], Dict[str, torch.Tensor]] = {}
    emb = {"embeddings.word_embeddings.weight": "model.embed_tokens.weight",
           "embeddings.position_embeddings.weight": "model.embed_positions.weight",
           "embeddings.token_type_embeddings.weight": "model.embed_token_types.weight",
           "embeddings.LayerNorm.weight": "model.embed_layernorm.weight",
           "embeddings.LayerNorm.bias": "model.embed_layernorm.bias"}
    for k, v in hf_sd.items():
        k = _strip(k, "bert.")
        if k in emb:
            out[emb[k]] = v
            continue
        m = re.match(r"encoder\.layer\.(\d+)\.(.*)", k)
        if not m:
            continue
        pre, rest = f"model.layers.{m.group(1)}.", m.group(2)
        mm = re.match(r"attention\.self\.(query|key|value)\.(weight|bias)", rest)
        if mm:
            parts.setdefault((pre, mm.group(2)), {})[mm.group(1)[0]] = v
            continue
        for a, b in {"attention.output.dense.": "self_attn.o_proj.", "attention.output.LayerNorm.": "input_layernorm.",
                     "intermediate.dense.": "mlp.up_proj.", "output.dense.": "mlp.down_proj.",
                     "output.LayerNorm.": "post_attention_layernorm."}.items():
            if rest.startswith(a):
                out[pre + b + rest[len(a):]] = v
                break
    for (pre, kind), p in parts.items():
        out[f"{pre}self_attn.qkv_proj.{kind}"] = torch.cat([p["q"], p["k"], p["v"]], 0)
    return out


def _convert_gptj(hf_sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    parts: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, v in hf_sd.items():
        k = _strip(k, "transformer.")
        m = re.match(r"h\.(\d+)\.(.*)", k)
        if m:
            pre, rest = f"model.layers.{m.group(1)}.", m.group(2)
            mm = re.match(r"attn\.([qkv])_proj\.weight", rest)
            if mm:
                parts.setdefault(pre, {})[mm.group(1)] = v
                continue
            for a, b in {"attn.out_proj.": "self_attn.o_proj.", "ln_1.": "input_layernorm.", "mlp.fc_in.": "mlp.up_proj.",
                         "mlp.fc_out.": "mlp.down_proj."}.items():
                if rest.startswith(a):
                    out[pre + b + rest[len(a):]] = v
                    break
        elif k == "wte.weight":
            out["model.embed_tokens.weight"] = v
        elif k.startswith("ln_f."):
            out["model.norm." + k[5:]] = v
        elif k.startswith("lm_head."):
            out[k] = v
    for pre, p in parts.items():
        out[pre + "self_attn.qkv_proj.weight"] = torch.cat([p["q"], p["k"], p["v"]], 0)
    return out


def _convert_deepseek_v3(hf_sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> Dict[str, torch.Tensor]:
    """MLA projections keep their names; dense / shared-expert gate+up are fused; routed experts arrive either stacked
    (`mlp.experts.gate_up_proj [E, 2I, H]`, transformers >= 5) or one module per expert (released checkpoints)."""
    out: Dict[str, torch.Tensor] = {}
    gu: Dict[str, Dict[str, torch.Tensor]] = {}
    experts: Dict[str, Dict[int, Dict[str, torch.Tensor]]] = {}
    for k, v in hf_sd.items():
        m = re.match(r"(model\.layers\.\d+\.mlp\.)experts\.(\d+)\.(gate|up|down)_proj\.weight", k)
        if m:
            experts.setdefault(m.group(1), {}).setdefault(int(m.group(2)), {})[m.group(3)] = v
            continue
        m = re.match(r"(model\.layers\.\d+\.mlp\.(?:shared_experts\.)?)(gate|up)_proj\.weight", k)
        if m:
            gu.setdefault(m.group(1), {})[m.group(2)] = v
            continue
        if k.endswith("mlp.experts.gate_up_proj"):
            out[k[: -len("gate_up_proj")] + "w_up"] = v
        elif k.endswith("mlp.experts.down_proj"):
            out[k[: -len("down_proj")] + "w_down"] = v
        elif k.endswith("mlp.gate.weight"):
            out[k[: -len("gate.weight")] + "router.gate.weight"] = v
        elif k.endswith("mlp.gate.e_score_correction_bias"):
            out[k[: -len("gate.e_score_correction_bias")] + "router.e_score_correction_bias"] = v
        elif re.match(r"model\.layers\.(\d+)\.", k) and int(re.match(r"model\.layers\.(\d+)\.", k).group(1)) >= cfg.num_hidden_layers:
            continue                                    # multi-token-prediction layers of the release are not used
        else:
            out[k] = v
    for pre, parts in gu.items():
        out[pre + "gate_up_proj.weight"] = torch.cat([parts["gate"], parts["up"]], dim=0)
    for pre, ex in experts.items():
        ids = sorted(ex)
        out[pre + "experts.w_up"] = torch.stack([torch.cat([ex[i]["gate"], ex[i]["up"]], 0) for i in ids])
        out[pre + "experts.w_down"] = torch.stack([ex[i]["down"] for i in ids])
    return out


_FAMILY_CONVERTERS = {"deepseek_v3": _convert_deepseek_v3, "gptj": _convert_gptj, "opt": _convert_opt, "bloom": _convert_bloom, "falcon": _convert_falcon, "bert": _convert_bert}


def to_hf_state_dict(model, cfg: Optional[ModelConfig] = None) -> Dict[str, torch.Tensor]:
    """Inverse of `convert_hf_state_dict` for the llama-like families (un-fuses qkv / gate_up / experts)."""
    cfg = cfg or model.cfg
    if cfg.model_type == "deepseek_v3":
        out = {}
        for k, v in model.state_dict().items():
            v = v.detach()
            if k.endswith("mlp.experts.w_up"):
                out[k[: -len("w_up")] + "gate_up_proj"] = v
            elif k.endswith("mlp.experts.w_down"):
                out[k[: -len("w_down")] + "down_proj"] = v
            elif k.endswith("gate_up_proj.weight"):
                g, u = v.chunk(2, dim=0)
                out[k[: -len("gate_up_proj.weight")] + "gate_proj.weight"] = g
                out[k[: -len("gate_up_proj.weight")] + "up_proj.weight"] = u
            elif k.endswith("mlp.router.gate.weight"):
                out[k.replace("router.gate.weight", "gate.weight")] = v
            elif k.endswith("mlp.router.e_score_correction_bias"):
                out[k.replace("router.e_score_correction_bias", "gate.e_score_correction_bias")] = v
            else:
                out[k] = v
        return out
    assert cfg.model_type in _LLAMA_LIKE, "export is implemented for the llama-like families"
    q, kv = cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim
    out: Dict[str, torch.Tensor] = {}
    for k, v in model.state_dict().items():
        v = v.detach()
        if k.endswith(("qkv_proj.weight", "qkv_proj.bias")):
            pre, kind = k.rsplit("qkv_proj.", 1)
            a, b, c = v.split([q, kv, kv], dim=0)
            out[pre + "q_proj." + kind], out[pre + "k_proj." + kind], out[pre + "v_proj." + kind] = a, b, c
        elif k.endswith("gate_up_proj.weight"):
            pre = k[: -len("gate_up_proj.weight")]
            g, u = v.chunk(2, dim=0)
            out[pre + "gate_proj.weight"], out[pre + "up_proj.weight"] = g, u
        elif k.endswith("mlp.experts.w_up"):
            pre = k[: -len("mlp.experts.w_up")] + "block_sparse_moe.experts."
            for i in range(v.shape[0]):
                w1, w3 = v[i].chunk(2, dim=0)
                out[f"{pre}{i}.w1.weight"], out[f"{pre}{i}.w3.weight"] = w1, w3
        elif k.endswith("mlp.experts.w_down"):
            pre = k[: -len("mlp.experts.w_down")] + "block_sparse_moe.experts."
            for i in range(v.shape[0]):
                out[f"{pre}{i}.w2.weight"] = v[i]
        elif k.endswith("mlp.router.gate.weight"):
            out[k.replace("mlp.router.gate.weight", "block_sparse_moe.gate.weight")] = v
        else:
            out[k] = v
    return out


def load_hf_checkpoint(path: str, dtype: torch.dtype = torch.bfloat16, device: str = "cpu", strict: bool = False):
    """Build our model from `path/config.json` and load every weight shard found in `path`."""
    from . import build_model

    with open(os.path.join(path, "config.json")) as f:
        cfg = config_from_hf(json.load(f))
    model = build_model(cfg).to(dtype)
    merged: Dict[str, torch.Tensor] = {}
    for shard in iter_hf_shards(path):
        merged.update(shard)        # q/k/v of one layer may sit in different shards -> convert once at the end
    sd = convert_hf_state_dict(merged, cfg)
    missing, unexpected = model.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, strict=False)
    missing = [m for m in missing if not (cfg.tie_word_embeddings and m == "lm_head.weight")]
    if strict and (missing or unexpected):
        raise RuntimeError(f"load_hf_checkpoint: missing={missing[:8]} unexpected={unexpected[:8]}")
    model._pretrained_path = path
    return model.to(device)


def is_hf_model(model) -> bool:
    """A `transformers.PreTrainedModel` instance (duck-typed: no hard dependency on transformers)."""
    mod = type(model).__module__ or ""
    return mod.startswith("transformers.") and hasattr(model, "config") and hasattr(model.config, "to_dict")


def from_hf_model(hf_model, dtype: Optional[torch.dtype] = None):
    """Convert an instantiated Hugging Face model into the equivalent colossalai_b200 model (same weights).

    This is what lets reference-style user code (`model = LlamaForCausalLM.from_pretrained(...)`;
    `booster.boost(model, ...)`) run unchanged: `Booster.boost` calls this for HF instances.  Decoder families go
    through `convert_hf_state_dict`; ViT / T5 / Whisper through `hf_io_encdec`."""
    hf_cfg = hf_model.config.to_dict()
    mt = hf_cfg.get("model_type", "")
    dtype = dtype or next(hf_model.parameters()).dtype
    sd = hf_model.state_dict()
    if mt in ("vit", "t5", "whisper"):
        from . import hf_io_encdec as ed
        from .t5 import T5EncoderModel, T5ForConditionalGeneration, T5Model
        from .vit import ViTForImageClassification, ViTModel
        from .whisper import WhisperForConditionalGeneration, WhisperModel

        name = type(hf_model).__name__
        if mt == "vit":
            cfg = ed.vit_config_from_hf(hf_cfg)
            cls = ViTForImageClassification if "Classification" in name else ViTModel
            if cls is ViTModel:
                sd = {("vit." + k if not k.startswith("vit.") else k): v for k, v in sd.items()}
                ours = cls(cfg, add_pooling_layer=any(k.startswith("vit.pooler") for k in sd))
                conv = {k[len("vit."):]: v for k, v in ed.convert_vit(sd).items()}
                ours.load_state_dict(conv, strict=False)
                return ours.to(dtype)
            ours = cls(cfg)
        elif mt == "t5":
            cfg = ed.t5_config_from_hf(hf_cfg)
            cls = T5ForConditionalGeneration if "ConditionalGeneration" in name else (
                T5EncoderModel if "Encoder" in name else T5Model)
            ours = cls(cfg)
        else:
            cfg = ed.whisper_config_from_hf(hf_cfg)
            if "ConditionalGeneration" in name:
                ours = WhisperForConditionalGeneration(cfg)
            else:
                ours = WhisperModel(cfg)
                sd = {k[len("model."):] if k.startswith("model.") else k: v for k, v in sd.items()}
        return ed.load_hf_encdec(ours, sd, strict=False).to(dtype)
    from . import build_model

    cfg = config_from_hf(hf_cfg)
    ours = build_model(cfg)
    conv = convert_hf_state_dict(sd, cfg)
    missing, unexpected = ours.load_state_dict(conv, strict=False)
    bad = [m for m in missing if not (m == "lm_head.weight" and cfg.tie_word_embeddings) and not m.endswith("norm.bias")
           and not m.endswith("layernorm.bias")]
    if bad:
        raise RuntimeError(f"from_hf_model({type(hf_model).__name__}): unmapped parameters {bad[:6]}")
    return ours.to(dtype)


def hf_config_dict(cfg: ModelConfig) -> dict:
    """`config.json` content for the llama-like families (enough for `AutoModelForCausalLM.from_pretrained`)."""
    if cfg.model_type == "deepseek_v3":
        m = cfg.moe
        return {"architectures": ["DeepseekV3ForCausalLM"], "model_type": "deepseek_v3", "vocab_size": cfg.vocab_size,
                "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
                "moe_intermediate_size": m.moe_intermediate_size, "num_hidden_layers": cfg.num_hidden_layers,
                "num_attention_heads": cfg.num_attention_heads, "num_key_value_heads": cfg.num_attention_heads,
                "n_shared_experts": m.n_shared_experts, "n_routed_experts": m.num_experts,
                "routed_scaling_factor": m.routed_scaling_factor, "kv_lora_rank": cfg.kv_lora_rank,
                "q_lora_rank": cfg.q_lora_rank, "qk_rope_head_dim": cfg.qk_rope_head_dim, "v_head_dim": cfg.v_head_dim,
                "qk_nope_head_dim": cfg.qk_nope_head_dim, "n_group": m.n_group, "topk_group": m.topk_group,
                "num_experts_per_tok": m.top_k, "first_k_dense_replace": m.first_k_dense_replace,
                "norm_topk_prob": m.norm_topk_prob, "hidden_act": cfg.hidden_act,
                "max_position_embeddings": cfg.max_position_embeddings, "rms_norm_eps": cfg.norm_eps,
                "rope_theta": cfg.rope_theta, "rope_scaling": cfg.rope_scaling, "rope_interleave": cfg.rope_interleaved,
                "attention_bias": cfg.attention_bias, "tie_word_embeddings": cfg.tie_word_embeddings,
                "bos_token_id": cfg.bos_token_id, "eos_token_id": cfg.eos_token_id, "torch_dtype": "bfloat16"}
    assert cfg.model_type in ("llama", "mistral", "qwen2", "qwen3", "mixtral"), cfg.model_type
    arch = {"llama": "LlamaForCausalLM", "mistral": "MistralForCausalLM", "qwen2": "Qwen2ForCausalLM",
            "qwen3": "Qwen3ForCausalLM", "mixtral": "MixtralForCausalLM"}[cfg.model_type]
    d = {"architectures": [arch], "model_type": cfg.model_type, "vocab_size": cfg.vocab_size,
         "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
         "num_hidden_layers": cfg.num_hidden_layers, "num_attention_heads": cfg.num_attention_heads,
         "num_key_value_heads": cfg.num_key_value_heads, "head_dim": cfg.head_dim,
         "max_position_embeddings": cfg.max_position_embeddings, "rms_norm_eps": cfg.norm_eps,
         "rope_theta": cfg.rope_theta, "rope_scaling": cfg.rope_scaling, "hidden_act": cfg.hidden_act,
         "tie_word_embeddings": cfg.tie_word_embeddings, "attention_bias": cfg.attention_bias,
         "bos_token_id": cfg.bos_token_id, "eos_token_id": cfg.eos_token_id, "torch_dtype": "bfloat16"}
    if cfg.sliding_window:
        d["sliding_window"] = cfg.sliding_window
    if cfg.moe is not None:
        d.update(num_local_experts=cfg.moe.num_experts, num_experts_per_tok=cfg.moe.top_k)
    return d


def save_hf_checkpoint(model, path: str, max_shard_bytes: int = 5 * 2**30, safe_serialization: bool = True) -> None:
    """Write a Hugging Face style directory (`config.json`, `model-0000x-of-0000N.safetensors`, index) from a —
    possibly tensor-/expert-parallel sharded and wrapped — model: every rank takes part in gathering the shards,
    rank 0 writes.  `transformers.AutoModelForCausalLM.from_pretrained(path)` loads the result."""
    import torch.distributed as dist

    from ..interface import ModelWrapper
    from ..tensor.d_tensor import to_global

    inner = model.unwrap() if isinstance(model, ModelWrapper) else model
    cfg = inner.cfg
    full: Dict[str, torch.Tensor] = {}
    for name, p in inner.named_parameters():
        t = to_global(p).detach()
        if name.endswith(("embed_tokens.weight", "lm_head.weight")) and t.shape[0] > cfg.vocab_size:
            t = t[: cfg.vocab_size]                       # drop the vocab padding added for tensor parallelism
        full[name] = t.cpu()
    if cfg.tie_word_embeddings and "lm_head.weight" not in full:
        full["lm_head.weight"] = full["model.embed_tokens.weight"]
    if dist.is_initialized() and dist.get_rank() != 0:
        dist.barrier()
        return

    class _Holder:
        pass

    h = _Holder()
    h.cfg = cfg
    h.state_dict = lambda: full
    sd = to_hf_state_dict(h, cfg)
    if cfg.tie_word_embeddings:
        sd.pop("lm_head.weight", None)
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        hf_cfg = hf_config_dict(cfg)
        hf_cfg["torch_dtype"] = str(next(iter(sd.values())).dtype).replace("torch.", "")
        json.dump(hf_cfg, f, indent=2)
    shards, cur, size = [], {}, 0
    for k, v in sd.items():
        nb = v.numel() * v.element_size()
        if cur and size + nb > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = v.contiguous()
        size += nb
    shards.append(cur)
    weight_map = {}
    ext = "safetensors" if safe_serialization else "bin"
    for i, sh in enumerate(shards):
        fname = f"model.{ext}" if len(shards) == 1 else f"model-{i + 1:05d}-of-{len(shards):05d}.{ext}"
        if safe_serialization:
            from safetensors.torch import save_file

            save_file(sh, os.path.join(path, fname), metadata={"format": "pt"})
        else:
            torch.save(sh, os.path.join(path, fname))
        weight_map.update({k: fname for k in sh})
    if len(shards) > 1:
        idx = "model.safetensors.index.json" if safe_serialization else "pytorch_model.bin.index.json"
        with open(os.path.join(path, idx), "w") as f:
            json.dump({"metadata": {"total_size": sum(v.numel() * v.element_size() for v in sd.values())},
                       "weight_map": weight_map}, f, indent=2)
    if dist.is_initialized():
        dist.barrier()
