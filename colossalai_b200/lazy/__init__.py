from .lazy_init import LazyInitContext, LazyTensor, copy_lazy_ops, is_lazy
from .pretrained import from_pretrained

__all__ = ["LazyInitContext", "LazyTensor", "copy_lazy_ops", "is_lazy", "from_pretrained"]
