from .layernorm import MixedFusedLayerNorm, MixedFusedRMSNorm
from .scaled_softmax import AttnMaskType, FusedScaleMaskSoftmax
from .utils import divide, get_tensor_parallel_mode

__all__ = ["MixedFusedLayerNorm", "MixedFusedRMSNorm", "FusedScaleMaskSoftmax", "AttnMaskType", "divide",
           "get_tensor_parallel_mode"]
