"""Supervised fine-tuning.  Parity: reference `coati/trainer/sft.py:1-248`."""
from __future__ import annotations

import torch

from ..models import GPTLMLoss, get_logits
from .base import SLTrainer


class SFTTrainer(SLTrainer):
    def __init__(self, model, booster, optim, lr_scheduler=None, max_epochs: int = 1, accumulation_steps: int = 1,
                 device=None) -> None:
        super().__init__(booster, max_epochs, model, optim, lr_scheduler, accumulation_steps, device)
        self.loss_fn = GPTLMLoss()

    def _train_step(self, batch):
        logits = get_logits(self.model, batch["input_ids"], batch.get("attention_mask"))
        return self.loss_fn(logits, batch["labels"]), {}
