from .base import BaseModel, Critic, RewardModel, disable_dropout, get_logits
from .generation import generate
from .loss import (DpoLoss, GPTLMLoss, KTOLoss, LogExpLoss, LogSigLoss, OddsRatioLoss, PolicyLoss, ValueLoss)
from .utils import (calc_action_log_probs, calc_masked_log_probs, compute_reward, log_probs_from_logits, masked_mean,
                    masked_whiten)

__all__ = ["BaseModel", "Critic", "RewardModel", "disable_dropout", "get_logits", "generate", "GPTLMLoss",
           "PolicyLoss", "ValueLoss", "DpoLoss", "LogSigLoss", "LogExpLoss", "OddsRatioLoss", "KTOLoss",
           "log_probs_from_logits", "calc_action_log_probs", "calc_masked_log_probs", "masked_mean", "masked_whiten",
           "compute_reward"]
