from .base import OLTrainer, SLTrainer
from .dpo import DPOTrainer
from .grpo import GRPOTrainer
from .kto import KTOTrainer
from .orpo import ORPOTrainer
from .ppo import PPOTrainer
from .rm import RewardModelTrainer
from .sft import SFTTrainer

__all__ = ["SLTrainer", "OLTrainer", "SFTTrainer", "RewardModelTrainer", "DPOTrainer", "ORPOTrainer", "KTOTrainer",
           "PPOTrainer", "GRPOTrainer"]
