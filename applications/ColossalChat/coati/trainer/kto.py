"""KTO on unpaired (completion, desirable?) data.  Parity: reference `coati/trainer/kto.py:1-355`."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..models import KTOLoss, calc_masked_log_probs, get_logits
from .base import SLTrainer


class KTOTrainer(SLTrainer):
    def __init__(self, actor: nn.Module, ref_model: nn.Module, booster, actor_optim, lr_scheduler=None,
                 beta: float = 0.1, desirable_weight: float = 1.0, undesirable_weight: float = 1.0, max_epochs: int = 1,
                 accumulation_steps: int = 1, device=None) -> None:
        super().__init__(booster, max_epochs, actor, actor_optim, lr_scheduler, accumulation_steps, device)
        self.ref_model = ref_model.eval()
        for p in ref_model.parameters():
            p.requires_grad_(False)
        self.loss_fn = KTOLoss(beta, desirable_weight, undesirable_weight)

    @staticmethod
    def _seq_logp(model, ids, am, lm):
        return calc_masked_log_probs(get_logits(model, ids, am), ids, lm).sum(-1)

    def _train_step(self, batch):
        lab = batch["label"].bool()
        lp = self._seq_logp(self.model, batch["input_ids"], batch["attention_mask"], batch["loss_mask"])
        kl = self._seq_logp(self.model, batch["kl_input_ids"], batch["kl_attention_mask"], batch["kl_loss_mask"])
        with torch.no_grad():
            rlp = self._seq_logp(self.ref_model, batch["input_ids"], batch["attention_mask"], batch["loss_mask"])
            rkl = self._seq_logp(self.ref_model, batch["kl_input_ids"], batch["kl_attention_mask"],
                                 batch["kl_loss_mask"])
        loss, cw, rw, z0 = self.loss_fn(lp[lab], lp[~lab], kl, rlp[lab], rlp[~lab], rkl)
        return loss, {"kl": float(z0), "chosen_reward": float(cw.mean()) if cw.numel() else 0.0,
                      "rejected_reward": float(rw.mean()) if rw.numel() else 0.0}
