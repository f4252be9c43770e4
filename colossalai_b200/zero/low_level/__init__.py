from .low_level_optim import LowLevelZeroOptimizer

__all__ = ["LowLevelZeroOptimizer"]
