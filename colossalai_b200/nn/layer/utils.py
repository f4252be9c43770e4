"""Parity: reference `colossalai/nn/layer/utils.py`."""


def divide(numerator: int, denominator: int) -> int:
    assert denominator != 0, "denominator can not be zero"
    assert numerator % denominator == 0, f"{numerator} is not divisible by {denominator}"
    return numerator // denominator


def get_tensor_parallel_mode():
    import os

    return os.environ.get("TENSOR_PARALLEL_MODE", None)
