from .weight_grad_store import WeightGradStore

__all__ = ["WeightGradStore"]
