from .lazy_init import LazyInitContext, LazyTensor, copy_lazy_ops, is_lazy

__all__ = ["LazyInitContext", "LazyTensor", "copy_lazy_ops", "is_lazy"]
