from .chunk import Chunk, ChunkManager, TensorState, init_chunk_manager, search_chunk_configuration
from .gemini_ddp import GeminiDDP
from .gemini_mgr import GeminiManager
from .gemini_optimizer import GeminiAdamOptimizer, GeminiOptimizer
from .placement_policy import AutoPlacementPolicy, StaticPlacementPolicy

__all__ = ["GeminiManager", "TensorState", "Chunk", "ChunkManager", "search_chunk_configuration", "init_chunk_manager",
           "GeminiDDP", "GeminiOptimizer", "GeminiAdamOptimizer", "StaticPlacementPolicy", "AutoPlacementPolicy"]
