"""Hugging Face checkpoint mapping for the encoder / encoder-decoder / vision families (ViT, T5, Whisper).

Our blocks fuse q/k/v of self-attention into `qkv_proj` and k/v of cross-attention into `kv_proj`; T5's gated FFN
(`wi_0`, `wi_1`) becomes `gate_up_proj`; T5 keeps one relative-position table per stack (HF stores it in block 0);
Whisper's bias-free `k_proj` gets a zero bias slice.  `load_hf_encdec(model, hf_state_dict)` fills a constructed model.
"""
from __future__ import annotations

import re
from typing import Dict, Optional

import torch

from .t5 import T5Config
from .vit import ViTConfig
from .whisper import WhisperConfig

__all__ = ["vit_config_from_hf", "t5_config_from_hf", "whisper_config_from_hf", "convert_vit", "convert_t5",
           "convert_whisper", "load_hf_encdec"]


def vit_config_from_hf(hf: dict) -> ViTConfig:
    return ViTConfig(image_size=hf["image_size"], patch_size=hf["patch_size"], num_channels=hf.get("num_channels", 3),
                     hidden_size=hf["hidden_size"], num_hidden_layers=hf["num_hidden_layers"],
                     num_attention_heads=hf["num_attention_heads"], intermediate_size=hf["intermediate_size"],
                     hidden_act=hf.get("hidden_act", "gelu"), layer_norm_eps=hf.get("layer_norm_eps", 1e-12),
                     qkv_bias=hf.get("qkv_bias", True), num_labels=len(hf.get("id2label", {})) or hf.get("num_labels", 2),
                     encoder_stride=hf.get("encoder_stride", 16))


def t5_config_from_hf(hf: dict) -> T5Config:
    return T5Config(vocab_size=hf["vocab_size"], d_model=hf["d_model"], d_kv=hf["d_kv"], d_ff=hf["d_ff"],
                    num_layers=hf["num_layers"], num_decoder_layers=hf.get("num_decoder_layers"),
                    num_heads=hf["num_heads"], relative_attention_num_buckets=hf.get("relative_attention_num_buckets", 32),
                    relative_attention_max_distance=hf.get("relative_attention_max_distance", 128),
                    layer_norm_epsilon=hf.get("layer_norm_epsilon", 1e-6),
                    feed_forward_proj=hf.get("feed_forward_proj", "relu"),
                    tie_word_embeddings=hf.get("tie_word_embeddings", True), pad_token_id=hf.get("pad_token_id", 0),
                    eos_token_id=hf.get("eos_token_id", 1), decoder_start_token_id=hf.get("decoder_start_token_id", 0))


def whisper_config_from_hf(hf: dict) -> WhisperConfig:
    return WhisperConfig(vocab_size=hf["vocab_size"], num_mel_bins=hf["num_mel_bins"], d_model=hf["d_model"],
                         encoder_layers=hf["encoder_layers"], decoder_layers=hf["decoder_layers"],
                         encoder_attention_heads=hf["encoder_attention_heads"],
                         decoder_attention_heads=hf["decoder_attention_heads"], encoder_ffn_dim=hf["encoder_ffn_dim"],
                         decoder_ffn_dim=hf["decoder_ffn_dim"], max_source_positions=hf["max_source_positions"],
                         max_target_positions=hf["max_target_positions"],
                         activation_function=hf.get("activation_function", "gelu"), pad_token_id=hf.get("pad_token_id", 50256),
                         eos_token_id=hf.get("eos_token_id", 50256),
                         decoder_start_token_id=hf.get("decoder_start_token_id", 50257))


def _cat(parts: Dict[str, torch.Tensor], order: str, like: Optional[torch.Tensor] = None) -> torch.Tensor:
    ts = []
    for o in order:
        t = parts.get(o)
        if t is None:                       # missing bias (Whisper k_proj): zeros of the right width
            t = torch.zeros_like(like)
        ts.append(t)
    return torch.cat(ts, 0)


def convert_vit(hf_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    qkv: Dict = {}
    for k, v in hf_sd.items():
        if k.startswith("vit.embeddings.patch_embeddings.projection."):
            out["vit.embeddings.patch_embeddings." + k.rsplit(".", 1)[1]] = v
            continue
        m = re.match(r"vit\.encoder\.layer\.(\d+)\.(.*)", k)
        if not m:
            out[k] = v                       # cls_token, position_embeddings, layernorm, classifier, pooler
            continue
        pre, rest = f"vit.layers.{m.group(1)}.", m.group(2)
        mm = re.match(r"attention\.attention\.(query|key|value)\.(weight|bias)", rest)
        if mm:
            qkv.setdefault((pre, mm.group(2)), {})[mm.group(1)[0]] = v
            continue
        for a, b in {"attention.output.dense.": "self_attn.o_proj.", "intermediate.dense.": "mlp.up_proj.",
                     "output.dense.": "mlp.down_proj.", "layernorm_before.": "norm1.", "layernorm_after.": "norm2."}.items():
            if rest.startswith(a):
                out[pre + b + rest[len(a):]] = v
                break
    for (pre, kind), p in qkv.items():
        out[f"{pre}self_attn.qkv_proj.{kind}"] = _cat(p, "qkv")
    if "vit.pooler.dense.weight" in out:
        out["vit.pooler.weight"], out["vit.pooler.bias"] = out.pop("vit.pooler.dense.weight"), out.pop("vit.pooler.dense.bias")
    return out


def convert_t5(hf_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    sa: Dict = {}
    ca: Dict = {}
    gu: Dict = {}
    for k, v in hf_sd.items():
        if k in ("shared.weight", "lm_head.weight") or k.startswith("classifier."):
            out[k] = v
            continue
        m = re.match(r"(encoder|decoder)\.(.*)", k)
        if not m:
            continue
        stack, rest = m.group(1), m.group(2)
        if rest == "embed_tokens.weight":
            continue
        if rest.startswith("final_layer_norm."):
            out[f"{stack}.final_layer_norm.weight"] = v
            continue
        mb = re.match(r"block\.(\d+)\.layer\.(\d+)\.(.*)", rest)
        i, j, tail = int(mb.group(1)), int(mb.group(2)), mb.group(3)
        pre = f"{stack}.block.{i}."
        ffn_idx = 2 if stack == "decoder" else 1
        if tail == "SelfAttention.relative_attention_bias.weight":
            out[f"{stack}.relative_attention_bias.weight"] = v
        elif tail.startswith("SelfAttention."):
            name = tail.split(".")[1]
            if name == "o":
                out[pre + "self_attn.o_proj.weight"] = v
            else:
                sa.setdefault(pre, {})[name] = v
        elif tail.startswith("EncDecAttention."):
            name = tail.split(".")[1]
            if name == "o":
                out[pre + "cross_attn.o_proj.weight"] = v
            elif name == "q":
                out[pre + "cross_attn.q_proj.weight"] = v
            else:
                ca.setdefault(pre, {})[name] = v
        elif tail == "layer_norm.weight":
            tgt = {0: "norm1", 1: "norm_cross" if stack == "decoder" else "norm2", 2: "norm2"}[j]
            out[pre + tgt + ".weight"] = v
        elif tail.startswith("DenseReluDense.") and j == ffn_idx:
            name = tail.split(".")[1]
            if name == "wo":
                out[pre + "mlp.down_proj.weight"] = v
            elif name == "wi":
                out[pre + "mlp.up_proj.weight"] = v
            else:
                gu.setdefault(pre, {})[name] = v
    for pre, p in sa.items():
        out[pre + "self_attn.qkv_proj.weight"] = _cat(p, "qkv")
    for pre, p in ca.items():
        out[pre + "cross_attn.kv_proj.weight"] = _cat(p, "kv")
    for pre, p in gu.items():
        out[pre + "mlp.gate_up_proj.weight"] = torch.cat([p["wi_0"], p["wi_1"]], 0)      # act(wi_0 x) * (wi_1 x)
    return out


def convert_whisper(hf_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    sa: Dict = {}
    ca: Dict = {}
    for k, v in hf_sd.items():
        m = re.match(r"(model\.)?(encoder|decoder)\.layers\.(\d+)\.(.*)", k)
        if not m:
            if k == "model.encoder.embed_positions.weight":
                out["__enc_pos__"] = v
            else:
                out[k] = v
            continue
        root, stack, i, rest = m.group(1) or "", m.group(2), m.group(3), m.group(4)
        pre = f"{root}{stack}.layers.{i}."
        mm = re.match(r"(self_attn|encoder_attn)\.([qkv])_proj\.(weight|bias)", rest)
        if mm:
            which, name, kind = mm.groups()
            if which == "self_attn":
                sa.setdefault((pre, kind), {})[name] = v
            elif name == "q":
                out[pre + f"cross_attn.q_proj.{kind}"] = v
            else:
                ca.setdefault((pre, kind), {})[name] = v
            continue
        for a, b in {"self_attn.out_proj.": "self_attn.o_proj.", "encoder_attn.out_proj.": "cross_attn.o_proj.",
                     "self_attn_layer_norm.": "norm1.", "encoder_attn_layer_norm.": "norm_cross.",
                     "final_layer_norm.": "norm2.", "fc1.": "mlp.up_proj.", "fc2.": "mlp.down_proj."}.items():
            if rest.startswith(a):
                out[pre + b + rest[len(a):]] = v
                break
    for (pre, kind), p in sa.items():
        out[f"{pre}self_attn.qkv_proj.{kind}"] = _cat(p, "qkv", like=p["q"])
    for (pre, kind), p in ca.items():
        out[f"{pre}cross_attn.kv_proj.{kind}"] = _cat(p, "kv", like=p["v"])
    return out


def load_hf_encdec(model: torch.nn.Module, hf_sd: Dict[str, torch.Tensor], strict: bool = True) -> torch.nn.Module:
    mt = getattr(model.cfg, "model_type", "")
    sd = {"vit": convert_vit, "t5": convert_t5, "whisper": convert_whisper}[mt](hf_sd)
    enc_pos = sd.pop("__enc_pos__", None)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if enc_pos is not None:                  # Whisper's fixed sinusoid table is a buffer on our side
        enc = model.model.encoder if hasattr(model, "model") else model.encoder
        enc.embed_positions = enc_pos.to(enc.embed_positions.dtype)
    tied = {"lm_head.weight", "proj_out.weight"}
    missing = [m for m in missing if m not in tied]
    if strict and (missing or [u for u in unexpected if "embed_positions" not in u]):
        raise RuntimeError(f"load_hf_encdec: missing={missing[:6]} unexpected={unexpected[:6]}")
    return model
