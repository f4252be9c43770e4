from .config import GenerationConfig, InferenceConfig
from .core import InferenceEngine

__all__ = ["InferenceConfig", "InferenceEngine", "GenerationConfig"]
