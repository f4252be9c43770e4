from .accelerator import Accelerator
from .booster import Booster
from .plugin import Plugin

__all__ = ["Accelerator", "Booster", "Plugin"]
