from .block_cache import CacheBlock
from .kvcache_manager import KVCacheManager, RPCKVCacheManager

__all__ = ["CacheBlock", "KVCacheManager", "RPCKVCacheManager"]
