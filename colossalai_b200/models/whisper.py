"""Whisper family: `WhisperModel`, `WhisperForConditionalGeneration`, `WhisperForAudioClassification`.

Encoder = two 1-D convs over log-mel frames (the second strides by 2) + fixed sinusoidal positions + pre-norm
blocks; decoder = token + learned position embeddings, blocks with causal self-attention and cross-attention over
the audio memory; the output projection is tied to the token embedding.

Parity: reference `colossalai/shardformer/policies/whisper.py:30-560` + `modeling/whisper.py:36-1000`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encdec import (BlockConfig, DecoderBlock, EncoderBlock, init_module_weights, make_norm, seq2seq_loss,
                     shift_tokens_right)

__all__ = ["WhisperConfig", "WhisperEncoder", "WhisperDecoder", "WhisperModel", "WhisperForConditionalGeneration",
           "WhisperForAudioClassification", "WHISPER_ZOO"]


@dataclass
class WhisperConfig:
    model_type: str = "whisper"
    vocab_size: int = 51865
    num_mel_bins: int = 80
    d_model: int = 384
    encoder_layers: int = 4
    decoder_layers: int = 4
    encoder_attention_heads: int = 6
    decoder_attention_heads: int = 6
    encoder_ffn_dim: int = 1536
    decoder_ffn_dim: int = 1536
    max_source_positions: int = 1500
    max_target_positions: int = 448
    activation_function: str = "gelu"
    dropout: float = 0.0
    init_std: float = 0.02
    pad_token_id: int = 50256
    eos_token_id: int = 50256
    decoder_start_token_id: int = 50257
    num_labels: int = 2
    classifier_proj_size: int = 256

    @property
    def hidden_size(self) -> int:
        return self.d_model

    def block(self, decoder: bool) -> BlockConfig:
        return BlockConfig(hidden_size=self.d_model,
                           num_heads=self.decoder_attention_heads if decoder else self.encoder_attention_heads,
                           ffn_dim=self.decoder_ffn_dim if decoder else self.encoder_ffn_dim,
                           act=self.activation_function, norm_eps=1e-5, pre_norm=True, dropout=self.dropout)

    def replace(self, **kw) -> "WhisperConfig":
        return replace(self, **kw)


WHISPER_ZOO: Dict[str, WhisperConfig] = {
    "whisper-tiny": WhisperConfig(),
    "whisper-base": WhisperConfig(d_model=512, encoder_layers=6, decoder_layers=6, encoder_attention_heads=8,
                                  decoder_attention_heads=8, encoder_ffn_dim=2048, decoder_ffn_dim=2048),
    "whisper-large-v3": WhisperConfig(vocab_size=51866, num_mel_bins=128, d_model=1280, encoder_layers=32,
                                      decoder_layers=32, encoder_attention_heads=20, decoder_attention_heads=20,
                                      encoder_ffn_dim=5120, decoder_ffn_dim=5120),
    "whisper-test": WhisperConfig(vocab_size=512, num_mel_bins=16, d_model=64, encoder_layers=2, decoder_layers=2,
                                  encoder_attention_heads=4, decoder_attention_heads=4, encoder_ffn_dim=128,
                                  decoder_ffn_dim=128, max_source_positions=32, max_target_positions=32,
                                  pad_token_id=0, eos_token_id=1, decoder_start_token_id=2),
}


def _sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


class WhisperEncoder(nn.Module):
    def __init__(self, cfg: WhisperConfig) -> None:
        super().__init__()
        self.cfg = cfg
        self.conv1 = nn.Conv1d(cfg.num_mel_bins, cfg.d_model, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(cfg.d_model, cfg.d_model, kernel_size=3, stride=2, padding=1)
        self.register_buffer("embed_positions", _sinusoids(cfg.max_source_positions, cfg.d_model), persistent=False)
        bc = cfg.block(False)
        self.layers = nn.ModuleList([EncoderBlock(bc) for _ in range(cfg.encoder_layers)])
        self.layer_norm = make_norm(bc)
        self.gradient_checkpointing = False

    def forward(self, input_features: torch.Tensor) -> torch.Tensor:
        """`input_features` [B, n_mels, frames] -> [B, frames/2, d_model]."""
        x = F.gelu(self.conv1(input_features.to(self.conv1.weight.dtype)))
        x = F.gelu(self.conv2(x)).transpose(1, 2)
        x = x + self.embed_positions[: x.shape[1]].to(x.dtype)
        for blk in self.layers:
            if self.gradient_checkpointing and self.training:
                x = torch.utils.checkpoint.checkpoint(blk, x, use_reentrant=False)
            else:
                x = blk(x)
        return self.layer_norm(x)


class WhisperDecoder(nn.Module):
    def __init__(self, cfg: WhisperConfig) -> None:
        super().__init__()
        self.cfg = cfg
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.d_model, padding_idx=cfg.pad_token_id)
        self.embed_positions = nn.Embedding(cfg.max_target_positions, cfg.d_model)
        bc = cfg.block(True)
        self.layers = nn.ModuleList([DecoderBlock(bc) for _ in range(cfg.decoder_layers)])
        self.layer_norm = make_norm(bc)
        self.gradient_checkpointing = False

    def forward(self, input_ids: torch.Tensor, memory: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                caches: Optional[List[dict]] = None) -> torch.Tensor:
        past = 0
        if caches is not None and caches[0].get("self") is not None:
            past = caches[0]["self"][0].shape[2]
        pos = torch.arange(past, past + input_ids.shape[1], device=input_ids.device)
        x = self.embed_tokens(input_ids) + self.embed_positions(pos)[None]
        for i, blk in enumerate(self.layers):
            args = dict(memory=memory, key_padding_mask=attention_mask, cache=None if caches is None else caches[i])
            if self.gradient_checkpointing and self.training:
                x = torch.utils.checkpoint.checkpoint(blk, x, use_reentrant=False, **args)
            else:
                x = blk(x, **args)
        return self.layer_norm(x)


class _WhisperBase(nn.Module):
    def __init__(self, cfg: WhisperConfig) -> None:
        super().__init__()
        self.cfg = self.config = cfg
        self.shard_config = None

    def gradient_checkpointing_enable(self, *a, **k) -> None:
        for m in self.modules():
            if isinstance(m, (WhisperEncoder, WhisperDecoder)):
                m.gradient_checkpointing = True


class WhisperModel(_WhisperBase):
    def __init__(self, config: Optional[WhisperConfig] = None, **kw) -> None:
        super().__init__(config or WhisperConfig(**kw))
        self.encoder = WhisperEncoder(self.cfg)
        self.decoder = WhisperDecoder(self.cfg)
        init_module_weights(self, self.cfg.init_std)

    def forward(self, input_features: Optional[torch.Tensor] = None, decoder_input_ids: Optional[torch.Tensor] = None,
                decoder_attention_mask: Optional[torch.Tensor] = None,
                encoder_outputs: Optional[torch.Tensor] = None, caches=None, **unused) -> Dict[str, torch.Tensor]:
        mem = encoder_outputs if encoder_outputs is not None else self.encoder(input_features)
        h = self.decoder(decoder_input_ids, mem, decoder_attention_mask, caches)
        return {"last_hidden_state": h, "encoder_last_hidden_state": mem}


class WhisperForConditionalGeneration(_WhisperBase):
    def __init__(self, config: Optional[WhisperConfig] = None, **kw) -> None:
        super().__init__(config or WhisperConfig(**kw))
        cfg = self.cfg
        self.model = WhisperModel(cfg)
        self.proj_out = nn.Linear(cfg.d_model, cfg.vocab_size, bias=False)
        init_module_weights(self.proj_out, cfg.init_std)
        self.proj_out.weight = self.model.decoder.embed_tokens.weight

    def get_output_embeddings(self):
        return self.proj_out

    def get_input_embeddings(self):
        return self.model.decoder.embed_tokens

    def forward(self, input_features: Optional[torch.Tensor] = None, decoder_input_ids: Optional[torch.Tensor] = None,
                decoder_attention_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                encoder_outputs: Optional[torch.Tensor] = None, caches=None, **unused) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        if decoder_input_ids is None and labels is not None:
            decoder_input_ids = shift_tokens_right(labels, cfg.pad_token_id, cfg.decoder_start_token_id)
        o = self.model(input_features=input_features, decoder_input_ids=decoder_input_ids,
                       decoder_attention_mask=decoder_attention_mask, encoder_outputs=encoder_outputs, caches=caches)
        logits = self.proj_out(o["last_hidden_state"])
        out = {"logits": logits, "encoder_last_hidden_state": o["encoder_last_hidden_state"]}
        if labels is not None:
            out["loss"] = seq2seq_loss(logits, labels, self.shard_config, cfg.vocab_size)
        return out

    @torch.no_grad()
    def generate(self, input_features: torch.Tensor, max_new_tokens: int = 20,
                 decoder_input_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        cfg = self.cfg
        mem = self.model.encoder(input_features)
        B = input_features.shape[0]
        caches = [dict() for _ in self.model.decoder.layers]
        cur = decoder_input_ids if decoder_input_ids is not None else torch.full(
            (B, 1), cfg.decoder_start_token_id, dtype=torch.long, device=input_features.device)
        out = [cur]
        done = torch.zeros(B, dtype=torch.bool, device=cur.device)
        for _ in range(max_new_tokens):
            h = self.model.decoder(cur, mem, None, caches)
            logits = self.proj_out(h[:, -1])
            if logits.shape[-1] < cfg.vocab_size and self.shard_config is not None:
                from ..parallel import comm

                logits = comm.all_gather(logits, -1, self.shard_config.tensor_parallel_process_group)
            cur = logits[..., :cfg.vocab_size].argmax(-1, keepdim=True).masked_fill(done[:, None], cfg.pad_token_id)
            out.append(cur)
            done |= cur.squeeze(1) == cfg.eos_token_id
            if bool(done.all()):
                break
        return torch.cat(out, dim=1)


class WhisperForAudioClassification(_WhisperBase):
    def __init__(self, config: Optional[WhisperConfig] = None, **kw) -> None:
        super().__init__(config or WhisperConfig(**kw))
        cfg = self.cfg
        self.encoder = WhisperEncoder(cfg)
        self.projector = nn.Linear(cfg.d_model, cfg.classifier_proj_size)
        self.classifier = nn.Linear(cfg.classifier_proj_size, cfg.num_labels)
        init_module_weights(self, cfg.init_std)

    def forward(self, input_features: torch.Tensor, labels: Optional[torch.Tensor] = None,
                **unused) -> Dict[str, torch.Tensor]:
        h = self.projector(self.encoder(input_features)).mean(dim=1)
        logits = self.classifier(h)
        out = {"logits": logits}
        if labels is not None:
            out["loss"] = F.cross_entropy(logits.float(), labels)
        return out
