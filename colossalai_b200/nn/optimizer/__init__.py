from .cpu_adam import CPUAdam
from .fused_adam import FusedAdam
from .hybrid_adam import HybridAdam
from .nvme_optimizer import NVMeOptimizer

__all__ = ["FusedAdam", "CPUAdam", "HybridAdam", "NVMeOptimizer", "FusedSGD", "FusedLAMB", "Lamb", "Lars", "CAME",
           "Adafactor", "GaLoreAdamW8bit", "DistributedLamb", "DistributedCAME", "DistributedAdaFactor",
           "DistGaloreAwamW", "cast_to_distributed"]

_LAZY = {"FusedSGD": "fused_sgd", "FusedLAMB": "fused_lamb", "Lamb": "lamb", "Lars": "lars", "CAME": "came",
         "Adafactor": "adafactor", "GaLoreAdamW8bit": "galore", "DistributedLamb": "distributed_lamb",
         "DistributedCAME": "distributed_came", "DistributedAdaFactor": "distributed_adafactor",
         "DistGaloreAwamW": "distributed_galore"}


def __getattr__(name):
    import importlib

    if name in _LAZY:
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)


def cast_to_distributed(optim):
    """Swap Lamb / CAME / Adafactor / GaLore for their TP/ZeRO-aware versions (same hyper-parameters)."""
    import importlib

    mapping = {"Lamb": "DistributedLamb", "CAME": "DistributedCAME", "Adafactor": "DistributedAdaFactor",
               "GaLoreAdamW8bit": "DistGaloreAwamW"}
    name = optim.__class__.__name__
    if name in mapping:
        try:
            cls = __getattr__(mapping[name])
        except Exception:
            return optim
        optim.__class__ = cls
        if hasattr(optim, "_post_cast"):
            optim._post_cast()
    return optim
