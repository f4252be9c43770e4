from .grad_scaler import BaseGradScaler, ConstantGradScaler, DynamicGradScaler
from .mixed_precision_mixin import BF16MixedPrecisionMixin, FP16MixedPrecisionMixin, MixedPrecisionMixin
from .mixed_precision_optimizer import MixedPrecisionOptimizer, NaiveFP16MixedPrecisionMixin

__all__ = ["BaseGradScaler", "ConstantGradScaler", "DynamicGradScaler", "BF16MixedPrecisionMixin",
           "FP16MixedPrecisionMixin", "MixedPrecisionMixin", "MixedPrecisionOptimizer", "NaiveFP16MixedPrecisionMixin"]
