from .naive_amp import (
    BF16MixedPrecisionMixin,
    ConstantGradScaler,
    DynamicGradScaler,
    FP16MixedPrecisionMixin,
    MixedPrecisionMixin,
    MixedPrecisionOptimizer,
)

__all__ = ["MixedPrecisionOptimizer", "DynamicGradScaler", "ConstantGradScaler", "MixedPrecisionMixin",
           "FP16MixedPrecisionMixin", "BF16MixedPrecisionMixin"]
