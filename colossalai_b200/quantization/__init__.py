from .bnb import quantize_model
from .bnb_config import BnbQuantizationConfig
from .fp8 import (all_gather_fp8, all_reduce_fp8, all_to_all_fp8, all_to_all_single_fp8, cast_from_fp8, cast_to_fp8,
                  linear_fp8, reduce_scatter_fp8)
from .fp8_hook import FP8Hook
from .gptq import GPTQ, QuantLinear, gptq_quantize_model
from .smoothquant import W8A8Linear, smooth_and_quantize_model

__all__ = ["BnbQuantizationConfig", "quantize_model", "FP8Hook", "cast_to_fp8", "cast_from_fp8", "linear_fp8",
           "all_reduce_fp8", "all_gather_fp8", "all_to_all_fp8", "all_to_all_single_fp8", "reduce_scatter_fp8", "GPTQ", "QuantLinear", "gptq_quantize_model", "W8A8Linear",
           "smooth_and_quantize_model"]
