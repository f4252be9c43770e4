"""Legacy 1-D tensor-parallel layers on the global parallel context.
Parity: reference `colossalai/legacy/nn/layer/parallel_1d/layers.py:1-1100` (`Linear1D`, `Linear1D_Col`,
`Linear1D_Row`, `Classifier1D`, `VocabParallelClassifier1D`, `Embedding1D`, `VocabParallelEmbedding1D`, `LayerNorm1D`,
`Dropout1D`).  They are the current shardformer layers bound to `ParallelMode.PARALLEL_1D`."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ....parallel import comm
from ....shardformer.layer import dropout as _dropout
from ....shardformer.layer import embedding as _emb
from ....shardformer.layer import linear as _lin
from ....shardformer.layer._operation import gather_forward_split_backward, reduce_forward
from ....shardformer.layer.normalization import FusedLayerNorm
from ...context import ParallelMode, global_context as gpc

__all__ = ["Linear1D", "Linear1D_Col", "Linear1D_Row", "Classifier1D", "VocabParallelClassifier1D", "Embedding1D",
           "VocabParallelEmbedding1D", "LayerNorm1D", "Dropout1D"]


def _pg():
    return gpc.get_group(ParallelMode.PARALLEL_1D) if gpc.is_initialized(ParallelMode.PARALLEL_1D) else None


class Linear1D_Col(_lin.Linear1D_Col):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, gather_output: bool = False,
                 skip_bias_add: bool = False, **kw) -> None:
        super().__init__(in_features, out_features, bias=bias, dtype=dtype, process_group=_pg(),
                         gather_output=gather_output, skip_bias_add=skip_bias_add, **kw)


class Linear1D_Row(_lin.Linear1D_Row):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None,
                 parallel_input: bool = True, skip_bias_add: bool = False, **kw) -> None:
        super().__init__(in_features, out_features, bias=bias, dtype=dtype, process_group=_pg(),
                         parallel_input=parallel_input, skip_bias_add=skip_bias_add, **kw)


class Linear1D(nn.Module):
    """Chooses column parallelism when the layer widens (out >= in... as the reference: `out_features < in_features`
    and no `gather_output` -> row) so consecutive linears alternate col/row without communication in between."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, dtype=None, gather_output: bool = False,
                 skip_bias_add: bool = False, **kw) -> None:
        super().__init__()
        parallel_input = kw.pop("parallel_input", None)
        row = (not gather_output) and out_features < in_features
        if row:
            self.layer = Linear1D_Row(in_features, out_features, bias=bias, dtype=dtype,
                                      parallel_input=True if parallel_input is None else parallel_input,
                                      skip_bias_add=skip_bias_add, **kw)
        else:
            self.layer = Linear1D_Col(in_features, out_features, bias=bias, dtype=dtype, gather_output=gather_output,
                                      skip_bias_add=skip_bias_add, **kw)

    @property
    def weight(self):
        return self.layer.weight

    @property
    def bias(self):
        return self.layer.bias

    def forward(self, x):
        return self.layer(x)


class Classifier1D(Linear1D_Row):
    """Classification head over a hidden-sharded input (row parallel, replicated logits)."""

    def __init__(self, in_features: int, num_classes: int, bias: bool = True, dtype=None, **kw) -> None:
        super().__init__(in_features, num_classes, bias=bias, dtype=dtype, parallel_input=kw.pop("parallel_input", True))


class VocabParallelClassifier1D(Linear1D_Col):
    """Class-dimension sharded head (logits stay sharded unless `gather_output`)."""

    def __init__(self, in_features: int, num_classes: int, bias: bool = True, dtype=None, gather_output: bool = False,
                 **kw) -> None:
        super().__init__(in_features, num_classes, bias=bias, dtype=dtype, gather_output=gather_output)


class Embedding1D(_emb.Embedding1D):
    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None, dtype=None,
                 **kw) -> None:
        super().__init__(num_embeddings, embedding_dim, padding_idx=padding_idx, dtype=dtype, process_group=_pg(), **kw)


class VocabParallelEmbedding1D(_emb.VocabParallelEmbedding1D):
    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None, dtype=None,
                 **kw) -> None:
        super().__init__(num_embeddings, embedding_dim, padding_idx=padding_idx, dtype=dtype, process_group=_pg(), **kw)


class LayerNorm1D(FusedLayerNorm):
    """Replicated LayerNorm (the reference's 1-D layernorm is not sharded either)."""

    def __init__(self, normalized_shape: int, eps: float = 1e-5, bias: bool = True, dtype=None) -> None:
        super().__init__(normalized_shape, eps=eps, bias=bias, dtype=dtype)


class Dropout1D(nn.Module):
    """Dropout on a tensor-parallel sharded activation: every rank draws from its own RNG stream."""

    def __init__(self, p: float = 0.5, inplace: bool = False) -> None:
        super().__init__()
        self.inner = _dropout.DropoutForParallelInput(p, inplace, process_group=_pg())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.inner(x)
