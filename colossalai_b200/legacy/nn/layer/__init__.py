from .parallel_1d import (Classifier1D, Dropout1D, Embedding1D, LayerNorm1D, Linear1D, Linear1D_Col, Linear1D_Row,
                          VocabParallelClassifier1D, VocabParallelEmbedding1D)
from .parallel_2d import Linear2D, split_2d
from .parallel_2p5d import Linear2p5D, split_2p5d
from .parallel_3d import Linear3D, split_3d_input
from .parallel_sequence import RingAV, RingQK, TransformerSelfAttentionRing

__all__ = ["Linear1D", "Linear1D_Col", "Linear1D_Row", "Classifier1D", "VocabParallelClassifier1D", "Embedding1D",
           "VocabParallelEmbedding1D", "LayerNorm1D", "Dropout1D", "Linear2D", "split_2d", "Linear2p5D", "split_2p5d",
           "Linear3D", "split_3d_input", "RingQK", "RingAV", "TransformerSelfAttentionRing"]
