"""Model-executor worker process for the RPC engine.

The reference drives one `rpyc` service per GPU (`colossalai/inference/executor/rpc_worker.py:44-308`: init_dist_env,
init_model, init_cache, execute_model_forward).  rpyc is not part of this image and is not needed: the workers live on
the same node as the scheduler, so we speak length-prefixed pickles over `multiprocessing.connection` (unix sockets /
loopback TCP).  Each worker owns one GPU, joins the TP process group over NCCL, holds its KV-cache shard and executes
forward steps; only rank 0 returns the sampled tokens.
"""
from __future__ import annotations

import os
import traceback
from multiprocessing.connection import Connection, Listener
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["RPCWorker", "serve_worker"]


class RPCWorker:
    """Methods callable from the scheduler process (`exposed_*` in the reference)."""

    def __init__(self) -> None:
        self.rank = 0
        self.world_size = 1
        self.engine_core = None

    # ---- lifecycle
    def init_dist_env(self, rank: int, world_size: int, master_address: str, master_port: int) -> str:
        from ... import initialize

        self.rank, self.world_size = rank, world_size
        initialize.launch(rank=rank, world_size=world_size, host=master_address, port=master_port,
                          backend="nccl" if torch.cuda.is_available() else "gloo", local_rank=rank)
        return "ok"

    def init_model(self, model_spec: Any, inference_config, model_policy=None) -> str:
        """`model_spec`: zoo name, ModelConfig or checkpoint path (an nn.Module cannot cross the process boundary
        cheaply; the reference has the same restriction and passes a path)."""
        from ...models import build_model
        from ...models.config import ModelConfig
        from ..core.llm_engine import LLMEngine

        torch.manual_seed(int(os.environ.get("CB200_INFER_SEED", "1234")))
        if isinstance(model_spec, str) and not os.path.isdir(model_spec):
            model_spec = build_model(model_spec)
        elif isinstance(model_spec, ModelConfig):
            model_spec = build_model(model_spec)
        self.engine_core = LLMEngine(model_spec, None, inference_config, model_policy=model_policy)
        return "ok"

    def init_cache(self, alloc_shape=None) -> Tuple[int, ...]:
        k, _ = self.engine_core.request_handler.get_kvcache()
        return tuple(k[0].shape)

    # ---- execution
    def execute_model_forward(self, input_token_ids: List[int], meta: Dict[str, Any], generation_config: Dict[str, Any]
                              ) -> Optional[List[int]]:
        """Runs one forward on the tokens/metadata chosen by the (remote) scheduler and samples."""
        from ..config import GenerationConfig
        from ..sampler import search_tokens

        eng = self.engine_core
        dev = eng.device
        ids = torch.tensor(input_token_ids, dtype=torch.long, device=dev)
        positions = eng.kv_runtime.set_step(torch.tensor(meta["block_tables"], dtype=torch.int32),
                                            torch.tensor(meta["sequence_lengths"], dtype=torch.int32),
                                            meta["is_prompts"], dev)
        with torch.inference_mode():
            logits = eng._model_forward(ids, positions)
            if meta["is_prompts"]:
                logits = logits[(eng.kv_runtime.cu_seqlens[1:] - 1).long()]
            logits = logits[:, : eng.model_config.vocab_size]
            gen = GenerationConfig(**{k: v for k, v in generation_config.items() if hasattr(GenerationConfig, k)
                                      or k in GenerationConfig.__dataclass_fields__})
            toks = search_tokens(gen, logits, meta["is_prompts"], batch_token_ids=meta.get("batch_token_ids"))
        if self.world_size > 1:
            dist.broadcast(toks, src=0)
        return toks.tolist() if self.rank == 0 else None

    def compute_only_for_test(self) -> float:
        x = torch.ones(4, device=self.engine_core.device if self.engine_core else "cpu")
        if self.world_size > 1:
            dist.all_reduce(x)
        return float(x.sum())

    def shutdown(self) -> str:
        if dist.is_initialized():
            dist.destroy_process_group()
        return "bye"


def serve_worker(address, authkey: bytes = b"cb200", ready_event=None) -> None:
    """Blocking server loop: one connection (the scheduler), messages are (method, args, kwargs)."""
    worker = RPCWorker()
    with Listener(address, authkey=authkey) as listener:
        if ready_event is not None:
            ready_event.set()
        conn: Connection = listener.accept()
        while True:
            try:
                method, args, kwargs = conn.recv()
            except EOFError:
                break
            try:
                res = getattr(worker, method)(*args, **kwargs)
                conn.send(("ok", res))
            except Exception as e:  # ship the traceback to the scheduler
                conn.send(("err", f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))
            if method == "shutdown":
                break
        conn.close()
