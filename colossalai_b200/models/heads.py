"""Task heads on top of the generic backbone (sequence / token classification, QA, masked LM, multiple choice).
Parity: the HF head classes covered by the reference's policies (`shardformer/policies/{llama,gpt2,bert,...}.py`)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..parallel import comm
from ..shardformer.layer._operation import gather_sp_output
from .config import ModelConfig
from .transformer import TransformerLMHeadModel, TransformerModel

__all__ = ["TransformerBackboneModel", "TransformerForSequenceClassification", "TransformerForTokenClassification",
           "TransformerForQuestionAnswering", "TransformerForMaskedLM", "TransformerForMultipleChoice"]


class _Base(nn.Module):
    def __init__(self, cfg: ModelConfig) -> None:
        super().__init__()
        self.cfg = self.config = cfg
        self.model = TransformerModel(cfg)
        self.shard_config = None

    def _init(self) -> None:
        std = self.cfg.initializer_range
        for m in self.modules():
            if isinstance(m, nn.Linear) and m.weight.device.type != "meta":
                nn.init.normal_(m.weight, std=std)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Embedding) and m.weight.device.type != "meta":
                nn.init.normal_(m.weight, std=std)

    def _hidden(self, input_ids, attention_mask=None, position_ids=None, token_type_ids=None, hidden_states=None,
                batch=None, seqlen=None):
        sc = self.shard_config
        B, S = input_ids.shape if input_ids is not None else (batch, seqlen)
        h = self.model(input_ids=input_ids, hidden_states=hidden_states, attention_mask=attention_mask,
                       position_ids=position_ids, token_type_ids=token_type_ids, batch=B, seqlen=S)
        sm = sc.pipeline_stage_manager if sc is not None else None
        if sm is not None and not sm.is_last_stage():
            return h, B, S, False
        if sc is not None and sc.sp_mode is not None and comm.group_size(sc.sp_group) > 1:
            if sc.sp_mode in ("split_gather", "ring"):
                h = gather_sp_output(h, sc.sp_group, sc.sp_mode, sp_dim=0)
            else:
                h = gather_sp_output(h.view(B, -1, h.shape[-1]), sc.sp_group, sc.sp_mode, sp_dim=1)
        return h.reshape(B, S, -1), B, S, True


class TransformerBackboneModel(_Base):
    """Bare backbone returning `last_hidden_state` [B, S, H]."""

    def __init__(self, cfg: ModelConfig) -> None:
        super().__init__(cfg)
        self._init()

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, token_type_ids=None,
                hidden_states=None, batch=None, seqlen=None, **unused) -> Dict[str, torch.Tensor]:
        h, B, S, last = self._hidden(input_ids, attention_mask, position_ids, token_type_ids, hidden_states, batch, seqlen)
        return {"last_hidden_state": h} if last else {"hidden_states": h}


class TransformerForSequenceClassification(_Base):
    def __init__(self, cfg: ModelConfig, num_labels: int = 2) -> None:
        super().__init__(cfg)
        self.num_labels = num_labels
        self.score = nn.Linear(cfg.hidden_size, num_labels, bias=not cfg.causal)
        self._init()

    def forward(self, input_ids=None, labels=None, attention_mask=None, position_ids=None, token_type_ids=None,
                hidden_states=None, batch=None, seqlen=None, **unused):
        h, B, S, last = self._hidden(input_ids, attention_mask, position_ids, token_type_ids, hidden_states, batch, seqlen)
        if not last:
            return {"hidden_states": h}
        if self.cfg.causal:   # last non-pad token
            if attention_mask is not None:
                idx = attention_mask.long().sum(-1) - 1
            else:
                idx = torch.full((B,), S - 1, device=h.device)
            pooled = h[torch.arange(B, device=h.device), idx]
        else:
            pooled = h[:, 0]
        logits = self.score(pooled)
        out = {"logits": logits}
        if labels is not None:
            out["loss"] = F.cross_entropy(logits.float(), labels) if self.num_labels > 1 else \
                F.mse_loss(logits.squeeze(-1).float(), labels.float())
        return out


class TransformerForTokenClassification(_Base):
    def __init__(self, cfg: ModelConfig, num_labels: int = 2) -> None:
        super().__init__(cfg)
        self.num_labels = num_labels
        self.classifier = nn.Linear(cfg.hidden_size, num_labels)
        self._init()

    def forward(self, input_ids=None, labels=None, attention_mask=None, position_ids=None, token_type_ids=None,
                hidden_states=None, batch=None, seqlen=None, **unused):
        h, B, S, last = self._hidden(input_ids, attention_mask, position_ids, token_type_ids, hidden_states, batch, seqlen)
        if not last:
            return {"hidden_states": h}
        logits = self.classifier(h)
        out = {"logits": logits}
        if labels is not None:
            out["loss"] = F.cross_entropy(logits.reshape(-1, self.num_labels).float(), labels.reshape(-1))
        return out


class TransformerForQuestionAnswering(_Base):
    def __init__(self, cfg: ModelConfig) -> None:
        super().__init__(cfg)
        self.qa_outputs = nn.Linear(cfg.hidden_size, 2)
        self._init()

    def forward(self, input_ids=None, start_positions=None, end_positions=None, attention_mask=None,
                position_ids=None, token_type_ids=None, hidden_states=None, batch=None, seqlen=None, **unused):
        h, B, S, last = self._hidden(input_ids, attention_mask, position_ids, token_type_ids, hidden_states, batch, seqlen)
        if not last:
            return {"hidden_states": h}
        start, end = self.qa_outputs(h).split(1, dim=-1)
        out = {"start_logits": start.squeeze(-1), "end_logits": end.squeeze(-1)}
        if start_positions is not None and end_positions is not None:
            out["loss"] = (F.cross_entropy(out["start_logits"].float(), start_positions) +
                           F.cross_entropy(out["end_logits"].float(), end_positions)) / 2
        return out


class TransformerForMaskedLM(TransformerLMHeadModel):
    """BERT-style MLM: same LM-head path, labels are NOT shifted (cfg.causal=False)."""


class TransformerForMultipleChoice(_Base):
    def __init__(self, cfg: ModelConfig) -> None:
        super().__init__(cfg)
        self.classifier = nn.Linear(cfg.hidden_size, 1)
        self._init()

    def forward(self, input_ids=None, labels=None, attention_mask=None, token_type_ids=None, **unused):
        B, C, S = input_ids.shape
        flat = lambda t: None if t is None else t.reshape(B * C, S)
        h, _, _, last = self._hidden(flat(input_ids), flat(attention_mask), None, flat(token_type_ids))
        if not last:
            return {"hidden_states": h}
        logits = self.classifier(h[:, 0]).view(B, C)
        out = {"logits": logits}
        if labels is not None:
            out["loss"] = F.cross_entropy(logits.float(), labels)
        return out
