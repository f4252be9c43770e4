from .model import AMPModelMixin, ModelWrapper, PeftUnwrapMixin
from .optimizer import DistributedOptim, OptimizerWrapper

__all__ = ["OptimizerWrapper", "ModelWrapper", "AMPModelMixin", "DistributedOptim", "PeftUnwrapMixin"]
