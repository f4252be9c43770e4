"""Seeding helper for tests (reference `colossalai/testing/random.py` `seed_all`)."""
import os
import random

import numpy as np
import torch

__all__ = ["seed_all"]


def seed_all(seed: int, cuda_deterministic: bool = False) -> None:
    """Seed python / numpy / torch (all devices).  `cuda_deterministic` trades cuDNN autotuning for run-to-run
    reproducibility; the default keeps autotuning on, which is what the performance tests want."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    cudnn = torch.backends.cudnn
    cudnn.deterministic, cudnn.benchmark = bool(cuda_deterministic), not cuda_deterministic
