"""Reference paths `colossalai/legacy/inference/quant/{gptq,smoothquant}` -> the maintained implementations."""
from ....quantization.gptq import GPTQ, CaiQuantLinear, QuantLinear, Quantizer, gptq_quantize_model
from ....quantization.smoothquant import (W8A8Linear, get_act_scales, smooth_and_quantize_model, smooth_ln_fcs)

__all__ = ["GPTQ", "CaiQuantLinear", "QuantLinear", "Quantizer", "gptq_quantize_model", "W8A8Linear",
           "get_act_scales", "smooth_and_quantize_model", "smooth_ln_fcs"]
