"""Automatic intra-op sharding plans.

Parity (capability): reference `colossalai/auto_parallel/tensor_shard` (`initialize_model`, `autoparallelize`,
strategy generation per node + ILP solver over resharding costs).  This implementation keeps the same interface on a
much smaller core: candidate strategies for every `nn.Linear` / `nn.Embedding` (replicate, column-shard, row-shard), a
communication cost from the mesh's alpha-beta model, a per-device memory budget, and an exact dynamic program over the
chain of layers (resharding cost between consecutive layers) instead of a general ILP.  The chosen plan is applied with
the Shardformer parallel layers.

Also here: `checkpoint` (Rotor activation-checkpoint solver, reference `auto_parallel/checkpoint`), `offload` (parameter
offload planner + runtime, reference `auto_parallel/offload`) and `autochunk` (activation chunking, reference
`colossalai/autochunk`)."""
from .autochunk import ChunkedModule, autochunk
from .checkpoint import Chain, CheckpointSolverRotor, apply_rotor_checkpointing
from .offload import AsynGreedySolver, SynGreedySolver, memory_optimize
from .solver import LayerStrategy, ShardingPlan, autoparallelize, initialize_model, solve_chain

__all__ = ["LayerStrategy", "ShardingPlan", "solve_chain", "initialize_model", "autoparallelize", "Chain",
           "CheckpointSolverRotor", "apply_rotor_checkpointing", "SynGreedySolver", "AsynGreedySolver",
           "memory_optimize", "ChunkedModule", "autochunk"]
