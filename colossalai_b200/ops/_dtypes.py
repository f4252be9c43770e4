import torch

DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
if hasattr(torch, "float8_e4m3fn"):
    DTYPE_CODE[torch.float8_e4m3fn] = 3
    DTYPE_CODE[torch.float8_e5m2] = 4


def code(dtype: torch.dtype) -> int:
    try:
        return DTYPE_CODE[dtype]
    except KeyError:
        raise TypeError(f"dtype {dtype} is not supported by the native kernels")
