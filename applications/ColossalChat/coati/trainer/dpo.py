"""Direct preference optimisation (and SimPO when no reference model is given and `gamma` / length normalisation are
set).  Parity: reference `coati/trainer/dpo.py:1-643`."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ..models import DpoLoss, calc_masked_log_probs, get_logits
from .base import SLTrainer


class DPOTrainer(SLTrainer):
    def __init__(self, actor: nn.Module, ref_model: Optional[nn.Module], booster, actor_optim, lr_scheduler=None,
                 beta: float = 0.1, gamma: float = 0.0, length_normalization: bool = False, max_epochs: int = 1,
                 accumulation_steps: int = 1, device=None) -> None:
        super().__init__(booster, max_epochs, actor, actor_optim, lr_scheduler, accumulation_steps, device)
        self.ref_model = ref_model
        if ref_model is not None:
            ref_model.eval()
            for p in ref_model.parameters():
                p.requires_grad_(False)
        self.loss_fn = DpoLoss(beta, gamma)
        self.length_normalization = length_normalization

    def _logps(self, model, batch, name):
        ids, am, lm = batch[f"{name}_input_ids"], batch[f"{name}_attention_mask"], batch[f"{name}_loss_mask"]
        return calc_masked_log_probs(get_logits(model, ids, am), ids, lm, self.length_normalization), lm[:, 1:].float()

    def _train_step(self, batch):
        lc, mc = self._logps(self.model, batch, "chosen")
        lr, mr = self._logps(self.model, batch, "rejected")
        rc = rr = None
        if self.ref_model is not None:
            with torch.no_grad():
                rc, _ = self._logps(self.ref_model, batch, "chosen")
                rr, _ = self._logps(self.ref_model, batch, "rejected")
        loss, cw, rw = self.loss_fn(lc, lr, rc, rr, mc, mr)
        return loss, {"chosen_reward": float(cw.mean()), "rejected_reward": float(rw.mean()),
                      "accuracy": float((cw > rw).float().mean())}
