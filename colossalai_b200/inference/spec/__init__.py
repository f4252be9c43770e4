from .drafter import Drafter
from .struct import DrafterOutput, GlideInput

__all__ = ["Drafter", "DrafterOutput", "GlideInput"]
