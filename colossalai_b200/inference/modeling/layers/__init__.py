from .distrifusion import (DistriConv2d, DistriKVExchange, PatchParallelContext, disable_patch_parallel,
                           enable_patch_parallel)

__all__ = ["DistriConv2d", "DistriKVExchange", "PatchParallelContext", "enable_patch_parallel",
           "disable_patch_parallel"]
