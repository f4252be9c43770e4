"""Automatic intra-op sharding plans.

Parity (capability): reference `colossalai/auto_parallel/tensor_shard` (`initialize_model`, `autoparallelize`,
strategy generation per node + ILP solver over resharding costs).  This implementation keeps the same interface on a
much smaller core: candidate strategies for every `nn.Linear` / `nn.Embedding` (replicate, column-shard, row-shard), a
communication cost from the mesh's alpha-beta model, a per-device memory budget, and an exact dynamic program over the
chain of layers (resharding cost between consecutive layers) instead of a general ILP.  The chosen plan is applied with
the Shardformer parallel layers.

`tensor_shard` is the graph-level counterpart (reference `auto_parallel/tensor_shard` + `auto_parallel/passes`): fx
trace, per-node strategy generation on 1-D / 2-D meshes, resharding costs, a 0/1 ILP (HiGHS) under a memory budget and
a runtime pass that shards parameters and inserts the collectives — `tensor_shard.initialize_model` /
`tensor_shard.autoparallelize`.

Also here: `checkpoint` (Rotor activation-checkpoint solver, reference `auto_parallel/checkpoint`), `offload` (parameter
offload planner + runtime, reference `auto_parallel/offload`) and `autochunk` (activation chunking, reference
`colossalai/autochunk`)."""
from . import tensor_shard
from .autochunk import ChunkedModule, autochunk
from .checkpoint import Chain, CheckpointSolverRotor, apply_rotor_checkpointing
from .offload import AsynGreedySolver, SynGreedySolver, memory_optimize
from .solver import LayerStrategy, ShardingPlan, autoparallelize, initialize_model, solve_chain

__all__ = ["LayerStrategy", "ShardingPlan", "solve_chain", "initialize_model", "autoparallelize", "Chain",
           "CheckpointSolverRotor", "apply_rotor_checkpointing", "SynGreedySolver", "AsynGreedySolver",
           "memory_optimize", "ChunkedModule", "autochunk", "tensor_shard"]
