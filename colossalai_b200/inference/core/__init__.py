from .engine import InferenceEngine
from .llm_engine import LLMEngine
from .request_handler import RequestHandler

__all__ = ["InferenceEngine", "LLMEngine", "RequestHandler"]
