from .kv_runtime import PagedKVRuntime

__all__ = ["PagedKVRuntime"]
