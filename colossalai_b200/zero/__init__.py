from .low_level import LowLevelZeroOptimizer

__all__ = ["LowLevelZeroOptimizer", "GeminiDDP", "GeminiOptimizer", "GeminiAdamOptimizer", "zero_model_wrapper",
           "zero_optim_wrapper"]


def __getattr__(name):
    if name in ("GeminiDDP", "GeminiOptimizer", "GeminiAdamOptimizer"):
        from . import gemini

        return getattr(gemini, name)
    if name in ("zero_model_wrapper", "zero_optim_wrapper"):
        from . import wrapper

        return getattr(wrapper, name)
    raise AttributeError(name)
