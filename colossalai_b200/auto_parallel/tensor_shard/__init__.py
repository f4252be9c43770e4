"""Graph-level intra-op auto parallelism: per-node sharding strategies over an fx graph, resharding costs between
them, an ILP for the cheapest consistent assignment under a memory budget, and a runtime pass that realises it
(reference `colossalai/auto_parallel/tensor_shard` + `auto_parallel/passes`)."""
from .initialize import (autoparallelize, build_strategy_constructor, extract_meta_args_from_dataloader,
                         initialize_device_mesh, initialize_model, shape_prop, solution_summary, solve_solution,
                         transform_to_sharded_model)
from .node_handler import (HandlerContext, generate_strategies, register_function_handler, register_module_handler,
                           reshape_dim_map)
from .runtime import ParallelEmbedding, ParallelLinear, reduce_bwd, reduce_fwd, reshard, runtime_apply_pass
from .sharding_strategy import ShardingStrategy, StrategiesVector, enumerate_specs, spec_str
from .solver import CostGraph, Solver, SolverOptions, StrategiesConstructor, resharding_cost, resharding_steps

__all__ = ["autoparallelize", "initialize_model", "initialize_device_mesh", "build_strategy_constructor",
           "solve_solution", "transform_to_sharded_model", "extract_meta_args_from_dataloader", "shape_prop",
           "solution_summary", "HandlerContext", "generate_strategies", "register_module_handler",
           "register_function_handler", "reshape_dim_map", "ParallelLinear", "ParallelEmbedding", "reshard",
           "reduce_fwd", "reduce_bwd", "runtime_apply_pass", "ShardingStrategy", "StrategiesVector", "enumerate_specs",
           "spec_str", "CostGraph", "Solver", "SolverOptions", "StrategiesConstructor", "resharding_cost",
           "resharding_steps"]
