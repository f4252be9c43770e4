from ._operation import all_to_all_comm, get_comm_backend, set_comm_backend
from .attn import AttnMaskType, ColoAttention, RingAttention
from .dropout import DropoutForParallelInput, DropoutForReplicatedInput
from .embedding import Embedding1D, PaddingEmbedding, VocabParallelEmbedding1D
from .linear import Linear1D_Col, Linear1D_Row, LinearWithGradAccum, PaddingLMHead, VocabParallelLMHead1D
from .loss import cross_entropy_1d, dist_cross_entropy, dist_log_prob, dist_log_prob_1d
from .normalization import BaseLayerNorm, FusedLayerNorm, FusedRMSNorm, LayerNorm, RMSNorm
from .parallel_module import PaddingParallelModule, ParallelModule
from .qkv_fused_linear import (
    FusedLinear,
    FusedLinear1D_Col,
    FusedLinear1D_Row,
    GPT2FusedLinearConv1D,
    GPT2FusedLinearConv1D_Col,
    GPT2FusedLinearConv1D_Row,
)
from .utils import Randomizer, RingComm, SeqParallelUtils

__all__ = [
    "all_to_all_comm", "get_comm_backend", "set_comm_backend", "AttnMaskType", "ColoAttention", "RingAttention",
    "DropoutForParallelInput", "DropoutForReplicatedInput", "Embedding1D", "PaddingEmbedding",
    "VocabParallelEmbedding1D", "Linear1D_Col", "Linear1D_Row", "LinearWithGradAccum", "PaddingLMHead",
    "VocabParallelLMHead1D", "cross_entropy_1d", "dist_cross_entropy", "dist_log_prob", "dist_log_prob_1d",
    "BaseLayerNorm", "FusedLayerNorm", "FusedRMSNorm", "LayerNorm", "RMSNorm", "PaddingParallelModule",
    "ParallelModule", "FusedLinear", "FusedLinear1D_Col", "FusedLinear1D_Row", "GPT2FusedLinearConv1D",
    "GPT2FusedLinearConv1D_Col", "GPT2FusedLinearConv1D_Row", "Randomizer", "RingComm", "SeqParallelUtils",
]
