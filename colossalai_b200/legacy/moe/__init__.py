"""The legacy capacity-based MoE stack (reference `colossalai/legacy/moe`): `Top1Router` / `Top2Router` / `TopKRouter`
(capacity factor, auxiliary load-balancing loss, optional noisy gating), expert-parallel `MLPExperts`,
`SparseMLP` layer built on the `MoeDispatch` / `MoeCombine` kernels and the all-to-all, `LoadBalancer` (swaps experts
between ranks according to observed load) and the `MOE_MANAGER` singleton.

Parity: `legacy/moe/layer/{routers.py:1-470, experts.py:1-160, layers.py:1-400}`, `legacy/moe/load_balance.py:1-440`,
`legacy/moe/manager.py:1-160`, `legacy/moe/utils.py`."""
from .experts import MLPExperts
from .layers import SparseMLP
from .load_balance import LoadBalancer
from .manager import MOE_MANAGER, MoEManager
from .routers import MoeRouter, Top1Router, Top2Router, TopKRouter, get_router_cls
from .utils import NormalNoiseGenerator, UniformNoiseGenerator, get_noise_generator

__all__ = ["MoeRouter", "Top1Router", "Top2Router", "TopKRouter", "get_router_cls", "MLPExperts", "SparseMLP",
           "LoadBalancer", "MOE_MANAGER", "MoEManager", "NormalNoiseGenerator", "UniformNoiseGenerator",
           "get_noise_generator"]
