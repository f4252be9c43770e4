from .plugin_base import Plugin
from .hybrid_parallel_plugin import HybridParallelPlugin

__all__ = ["Plugin", "HybridParallelPlugin", "TorchDDPPlugin", "TorchFSDPPlugin", "LowLevelZeroPlugin",
           "GeminiPlugin", "MoeHybridParallelPlugin"]


def __getattr__(name):
    import importlib

    table = {"TorchDDPPlugin": "torch_ddp_plugin", "TorchFSDPPlugin": "torch_fsdp_plugin",
             "LowLevelZeroPlugin": "low_level_zero_plugin", "GeminiPlugin": "gemini_plugin",
             "MoeHybridParallelPlugin": "moe_hybrid_parallel_plugin"}
    if name in table:
        return getattr(importlib.import_module(f"{__name__}.{table[name]}"), name)
    raise AttributeError(name)
