"""Pairwise reward-model training.  Parity: reference `coati/trainer/rm.py:1-247`."""
from __future__ import annotations

import torch

from ..models import LogExpLoss, LogSigLoss
from .base import SLTrainer


class RewardModelTrainer(SLTrainer):
    def __init__(self, model, booster, optimizer, lr_scheduler=None, loss_fn: str = "log_sig", max_epochs: int = 1,
                 accumulation_steps: int = 1, device=None) -> None:
        super().__init__(booster, max_epochs, model, optimizer, lr_scheduler, accumulation_steps, device)
        self.loss_fn = LogSigLoss() if loss_fn == "log_sig" else LogExpLoss()

    def _train_step(self, batch):
        # one forward over [chosen; rejected] keeps both halves in the same kernels / same dropout state
        ids_c, ids_r = batch["chosen_input_ids"], batch["rejected_input_ids"]
        n = max(ids_c.shape[1], ids_r.shape[1])
        pad = lambda t: torch.nn.functional.pad(t, (0, n - t.shape[1]))
        ids = torch.cat([pad(ids_c), pad(ids_r)], 0)
        mask = torch.cat([pad(batch["chosen_attention_mask"]), pad(batch["rejected_attention_mask"])], 0)
        r = self.model(ids, mask)
        rc, rr = r.chunk(2)
        loss = self.loss_fn(rc, rr)
        return loss, {"accuracy": float((rc > rr).float().mean()), "reward_margin": float((rc - rr).mean())}
