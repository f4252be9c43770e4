from .attention_backend import (AttentionBackend, AttentionMetaData, CudaAttentionBackend, ReferenceAttentionBackend,
                                get_attention_backend)
from .pre_attention_backend import (CudaPreAttentionBackend, PreAttentionBackend, ReferencePreAttentionBackend,
                                    get_pre_attention_backend)

__all__ = ["AttentionBackend", "AttentionMetaData", "CudaAttentionBackend", "ReferenceAttentionBackend",
           "get_attention_backend", "PreAttentionBackend", "CudaPreAttentionBackend", "ReferencePreAttentionBackend",
           "get_pre_attention_backend"]
