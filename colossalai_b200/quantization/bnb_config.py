"""Weight-only quantisation config (API of the reference's `BnbQuantizationConfig`, bnb_config.py:11-113)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

__all__ = ["BnbQuantizationConfig"]


@dataclass
class BnbQuantizationConfig:
    load_in_8bit: bool = False
    llm_int8_threshold: float = 6.0
    load_in_4bit: bool = False
    bnb_4bit_quant_type: str = "fp4"            # {"fp4", "nf4"}
    bnb_4bit_use_double_quant: bool = False
    bnb_4bit_compute_dtype: str = "fp16"        # {"fp32", "fp16", "bf16"}
    torch_dtype: Optional[torch.dtype] = None
    skip_modules: Optional[List[str]] = None
    keep_in_fp32_modules: Optional[List[str]] = None
    block_size: int = 64

    def __post_init__(self) -> None:
        if self.load_in_8bit and self.load_in_4bit:
            raise ValueError("load_in_8bit and load_in_4bit can't be both True")
        if not self.load_in_8bit and not self.load_in_4bit:
            raise ValueError("load_in_8bit and load_in_4bit can't be both False")
        if self.bnb_4bit_quant_type not in ("fp4", "nf4"):
            raise ValueError(f"bnb_4bit_quant_type must be in ['fp4','nf4'] but found {self.bnb_4bit_quant_type}")
        if isinstance(self.bnb_4bit_compute_dtype, str):
            m = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
            if self.bnb_4bit_compute_dtype not in m:
                raise ValueError("bnb_4bit_compute_dtype must be in ['fp32','fp16','bf16']")
            self.bnb_4bit_compute_dtype = m[self.bnb_4bit_compute_dtype]
        if self.skip_modules is not None and not isinstance(self.skip_modules, list):
            raise ValueError("skip_modules must be a list of strings")
        if self.torch_dtype is None:
            self.torch_dtype = self.bnb_4bit_compute_dtype if self.load_in_4bit else torch.float16
