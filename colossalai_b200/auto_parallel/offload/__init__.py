from .solver import AsynGreedySolver, OffloadPlan, SynGreedySolver, memory_optimize

__all__ = ["OffloadPlan", "SynGreedySolver", "AsynGreedySolver", "memory_optimize"]
