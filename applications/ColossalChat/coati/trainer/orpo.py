"""ORPO: SFT loss on the chosen answer + lambda * odds-ratio penalty, no reference model.
Parity: reference `coati/trainer/orpo.py:1-330`."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..models import OddsRatioLoss, calc_masked_log_probs, get_logits
from .base import SLTrainer


class ORPOTrainer(SLTrainer):
    def __init__(self, actor, booster, actor_optim, lr_scheduler=None, lam: float = 0.1, max_epochs: int = 1,
                 accumulation_steps: int = 1, device=None) -> None:
        super().__init__(booster, max_epochs, actor, actor_optim, lr_scheduler, accumulation_steps, device)
        self.lam, self.or_loss = lam, OddsRatioLoss()

    def _train_step(self, batch):
        ids_c, am_c, lm_c = batch["chosen_input_ids"], batch["chosen_attention_mask"], batch["chosen_loss_mask"]
        ids_r, am_r, lm_r = batch["rejected_input_ids"], batch["rejected_attention_mask"], batch["rejected_loss_mask"]
        logits_c = get_logits(self.model, ids_c, am_c)
        lc = calc_masked_log_probs(logits_c, ids_c, lm_c)
        lr = calc_masked_log_probs(get_logits(self.model, ids_r, am_r), ids_r, lm_r)
        labels = ids_c.masked_fill(lm_c == 0, -100)
        sft = F.cross_entropy(logits_c[:, :-1].reshape(-1, logits_c.shape[-1]).float(), labels[:, 1:].reshape(-1),
                              ignore_index=-100)
        ratio, log_odds = self.or_loss(lc, lr, lm_c[:, 1:].float(), lm_r[:, 1:].float())
        return sft + self.lam * ratio, {"sft_loss": float(sft), "log_odds_ratio": float(log_odds.mean())}
