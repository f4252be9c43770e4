"""Parity: reference `colossalai/quantization/fp8_config.py`."""
dynamic_kernel: bool = False
