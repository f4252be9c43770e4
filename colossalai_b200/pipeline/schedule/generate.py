"""Pipeline-parallel token generation.

Parity: reference `colossalai/pipeline/schedule/generate.py:34-441` (`GenerateSchedule`: micro-batches circulate through
the stages; the last stage turns hidden states into the next token and hands it back to the first stage) and its
`MicroBatchManager` bookkeeping.  The KV cache of every micro-batch lives on the stage that owns the layers
(`DenseKVCache`, the `kv_cache=` runtime of our generic transformer).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist
from torch.nn import Module

from ..p2p import PipelineP2PCommunication
from ..stage_manager import PipelineStageManager
from .base import PipelineSchedule

__all__ = ["GenerateSchedule", "MicroBatchManager", "DenseKVCache", "Status"]


class Status:
    PREFILL = 1
    GENERATE = 2
    DONE = 3


class DenseKVCache:
    """Append-only per-layer K/V store for ONE micro-batch of equal-length sequences (`kv_cache=` runtime)."""

    def __init__(self, batch: int) -> None:
        self.batch = batch
        self.k: Dict[int, torch.Tensor] = {}
        self.v: Dict[int, torch.Tensor] = {}

    def attend(self, layer_idx: int, q, k, v, meta, scale: float, alibi_slopes=None, sliding_window=None):
        from ... import ops

        B = self.batch
        kb = k.view(B, -1, *k.shape[1:])
        vb = v.view(B, -1, *v.shape[1:])
        if layer_idx in self.k:
            kb = torch.cat([self.k[layer_idx], kb], dim=1)
            vb = torch.cat([self.v[layer_idx], vb], dim=1)
        self.k[layer_idx], self.v[layer_idx] = kb, vb
        if sliding_window is not None and kb.shape[1] > sliding_window and alibi_slopes is None:
            alibi_slopes = torch.zeros(q.shape[1], device=q.device)          # band mask through the bias path
        if alibi_slopes is not None:
            # queries are the last Sq positions of the Sk cached ones: bias = slope * (key_pos - query_pos), causal
            Sq, Sk = q.shape[0] // B, kb.shape[1]
            qpos = torch.arange(Sk - Sq, Sk, device=q.device)
            rel = torch.arange(Sk, device=q.device)[None, :] - qpos[:, None]
            bias = rel.clamp(max=0).float()[None, None] * alibi_slopes.float()[None, :, None, None]
            hidden = rel > 0
            if sliding_window is not None:
                hidden = hidden | (rel <= -sliding_window)
            bias = bias.masked_fill(hidden[None, None], float("-inf")).to(q.dtype)
            return ops.attention(q, kb.reshape(-1, *k.shape[1:]), vb.reshape(-1, *v.shape[1:]), batch=B, causal=False,
                                 scale=scale, attn_mask=bias)
        return ops.attention(q, kb.reshape(-1, *k.shape[1:]), vb.reshape(-1, *v.shape[1:]), batch=B, causal=True,
                             scale=scale)

    @property
    def length(self) -> int:
        return next(iter(self.k.values())).shape[1] if self.k else 0


@dataclass
class _MicroBatch:
    input_ids: torch.Tensor
    max_new_tokens: int
    cache: DenseKVCache
    generated: List[torch.Tensor] = field(default_factory=list)
    prompt_len: int = 0

    @property
    def state(self) -> int:
        if not self.generated and self.cache.length == 0:
            return Status.PREFILL
        return Status.DONE if len(self.generated) >= self.max_new_tokens else Status.GENERATE


class MicroBatchManager:
    """Tracks which micro-batch a stage is working on and its generation state."""

    def __init__(self, stage: int, new_length: int, micro_batch_size: int, micro_batch_buffer_size: int) -> None:
        self.stage = stage
        self.new_length = new_length
        self.micro_batch_size = micro_batch_size
        self.buffer_size = micro_batch_buffer_size
        self.mbs: Dict[int, _MicroBatch] = {}
        self.idx = 0

    def add(self, input_ids: torch.Tensor) -> None:
        i = len(self.mbs)
        self.mbs[i] = _MicroBatch(input_ids, self.new_length, DenseKVCache(input_ids.shape[0]),
                                  prompt_len=input_ids.shape[1])

    def clear(self) -> None:
        self.mbs.clear()
        self.idx = 0

    @property
    def cur(self) -> _MicroBatch:
        return self.mbs[self.idx]

    def next(self) -> None:
        self.idx = (self.idx + 1) % max(len(self.mbs), 1)

    def is_done(self) -> bool:
        return all(len(m.generated) >= m.max_new_tokens for m in self.mbs.values())


class GenerateSchedule(PipelineSchedule):
    """Greedy generation over pipeline stages.  `generate_step(model, data_iter)` returns, on the LAST stage, a list of
    `[micro_batch, new_tokens]` tensors (one per micro-batch); other stages return `[]`."""

    def __init__(self, stage_manager: PipelineStageManager, mb_manager: MicroBatchManager, verbose: bool = False) -> None:
        super().__init__(stage_manager)
        self.comm = PipelineP2PCommunication(stage_manager)
        self.mb_manager = mb_manager
        self.verbose = verbose
        self.microbatch_size = mb_manager.micro_batch_size
        self.batch: Optional[Any] = None
        self._pending: List[Any] = []

    # ---- data
    def load_batch(self, data_iter: Iterable, device: Optional[torch.device] = None) -> None:
        batch = next(data_iter)
        ids = batch["input_ids"] if isinstance(batch, dict) else batch
        if device is not None:
            ids = ids.to(device)
        self.mb_manager.clear()
        for chunk in ids.split(self.microbatch_size, dim=0):
            self.mb_manager.add(chunk)

    # ---- one micro-batch through this stage
    def _stage_forward(self, model: Module, mb: _MicroBatch, incoming: Optional[torch.Tensor], new_token: Optional[torch.Tensor]):
        sm = self.stage_manager
        B = mb.input_ids.shape[0]
        prefill = mb.cache.length == 0
        S = mb.prompt_len if prefill else 1
        past = 0 if prefill else mb.cache.length
        pos = torch.arange(past, past + S, device=mb.input_ids.device).repeat(B)
        kwargs = dict(position_ids=pos.view(B, S), kv_cache=mb.cache, batch=B, seqlen=S)
        if sm.is_first_stage():
            ids = mb.input_ids if prefill else new_token.view(B, 1)
            return model(input_ids=ids, **kwargs)
        return model(hidden_states=incoming, **kwargs)

    @torch.no_grad()
    def generate_step(self, model: Module, data_iter: Iterable) -> List[torch.Tensor]:
        sm = self.stage_manager
        self.load_batch(data_iter, device=next(model.parameters()).device)
        mgr = self.mb_manager
        model.eval()
        n_mb = len(mgr.mbs)
        steps = mgr.new_length
        first, last = sm.is_first_stage(), sm.is_last_stage()
        single = sm.num_stages == 1
        for step in range(steps):
            for i in range(n_mb):
                mb = mgr.mbs[i]
                new_token = None
                if first and step > 0:
                    new_token = mb.generated[-1] if single else self._recv_token(mb)
                incoming = None
                if not first:
                    incoming = self.comm._communicate(None, None, sm.get_prev_rank()).wait()
                    if isinstance(incoming, dict):
                        incoming = incoming["hidden_states"]
                out = self._stage_forward(model, mb, incoming, new_token)
                if last:
                    logits = out["logits"]
                    B = mb.input_ids.shape[0]
                    tok = logits.view(B, -1, logits.shape[-1])[:, -1, : model.cfg.vocab_size].argmax(-1)
                    mb.generated.append(tok)
                    if not single and step < steps - 1:
                        self._send_token(tok)
                else:
                    self.comm._communicate({"hidden_states": out["hidden_states"]}, sm.get_next_rank(), None).wait()
        for h, _ in self._pending:
            h.wait()
        self._pending.clear()
        if last:
            return [torch.stack(mgr.mbs[i].generated, dim=1) for i in range(n_mb)]
        return []

    # the sampled token travels last stage -> first stage
    def _send_token(self, tok: torch.Tensor) -> None:
        # next of the last stage wraps to stage 0; asynchronous so the last stage keeps draining hidden states
        t = tok.contiguous()
        self._pending.append((dist.isend(t, dst=self.stage_manager.get_next_rank()), t))

    def _recv_token(self, mb: _MicroBatch) -> torch.Tensor:
        src = self.stage_manager.get_prev_rank()                                  # prev of stage 0 wraps to the last
        tok = torch.empty(mb.input_ids.shape[0], dtype=torch.long, device=mb.input_ids.device)
        dist.recv(tok, src=src)
        mb.generated.append(tok)          # the first stage mirrors the generated tokens
        return tok

    # training entry point of the base class is meaningless for generation
    def forward_backward_step(self, *args, **kwargs):
        raise NotImplementedError("GenerateSchedule only supports generate_step()")
