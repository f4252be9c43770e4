"""GeminiDDP: ZeRO-3 with chunked parameters and heterogeneous (HBM <-> pinned host) placement.

Parity: reference `colossalai/zero/gemini/gemini_ddp.py:56-989` + `gemini_hook.py:20-108`: parameters live in chunks
sharded over the zero group; a chunk is all-gathered just before the first module that needs it runs and released right
after; gradients are written into a gradient chunk which is reduce-scattered as soon as every tensor in it is ready and
then moved to the device that holds the matching optimizer shard; state-dicts are gathered chunk by chunk.
The just-in-time hooks are module-granular (identity autograd Functions around every parameter-owning module) — the
same pre-fwd / post-fwd / pre-bwd / post-bwd protocol the reference drives through `ColoParameter.__torch_function__`.
"""
from __future__ import annotations

import itertools
from collections import OrderedDict
from contextlib import nullcontext
from typing import Any, Callable, Dict, Iterator, List, Optional, Set, Tuple, Union

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.distributed import ProcessGroup

from ...accelerator import get_accelerator
from ...interface import ModelWrapper
from ...logging import get_dist_logger
from ...utils import is_ddp_ignored
from .chunk import Chunk, ChunkManager, TensorState, init_chunk_manager
from .gemini_mgr import GeminiManager
from .memory_tracer import MemStats

__all__ = ["GeminiDDP"]


class _PreBackward(torch.autograd.Function):
    """Identity on module OUTPUTS; its backward runs before the module's own backward."""

    @staticmethod
    def forward(ctx, owner, mod, *tensors):
        ctx.owner, ctx.mod = owner, mod
        return tensors if len(tensors) > 1 else tensors[0]

    @staticmethod
    def backward(ctx, *grads):
        ctx.owner._pre_backward(ctx.mod)
        return (None, None) + grads


class _PostBackward(torch.autograd.Function):
    """Identity on module INPUTS; its backward runs after the module's own backward."""

    @staticmethod
    def forward(ctx, owner, mod, *tensors):
        ctx.owner, ctx.mod = owner, mod
        return tensors if len(tensors) > 1 else tensors[0]

    @staticmethod
    def backward(ctx, *grads):
        ctx.owner._post_backward_module(ctx.mod)
        return (None, None) + grads


def _map_tensors(obj, fn):
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_tensors(o, fn) for o in obj)
    if isinstance(obj, dict):
        return {k: _map_tensors(v, fn) for k, v in obj.items()}
    return obj


class GeminiDDP(ModelWrapper):
    def __init__(self, module: nn.Module, chunk_config_dict: Optional[dict] = None,
                 chunk_init_device: torch.device = torch.device("cpu"), placement_policy: str = "static",
                 enable_gradient_accumulation: bool = False, max_prefetch: int = 0, shard_param_frac: float = 1.0,
                 offload_optim_frac: float = 0.0, offload_param_frac: float = 0.0,
                 warmup_non_model_data_ratio: float = 0.8, steady_cuda_cap_ratio: float = 0.9,
                 search_range_m: int = 32, hidden_dim: Optional[int] = None, min_chunk_size_m: float = 32,
                 pin_memory: bool = False, force_outputs_fp32: bool = False, strict_ddp_mode: bool = False,
                 scatter_after_inference: bool = True, mixed_precision: torch.dtype = torch.float16,
                 zero_group: Optional[ProcessGroup] = None, memstats: Optional[MemStats] = None,
                 master_weights: bool = True, extra_dp_group: Optional[ProcessGroup] = None, verbose: bool = False,
                 enable_async_reduce: bool = True, fp8_communication: bool = False, use_fp8: bool = False) -> None:
        assert mixed_precision in (torch.float16, torch.bfloat16, torch.float32)
        super().__init__(module)
        self.logger = get_dist_logger()
        self.zero_group = zero_group
        self.extra_dp_group = extra_dp_group
        self.mixed_precision = mixed_precision
        self.master_weights = master_weights
        self.force_outputs_fp32 = force_outputs_fp32
        self.scatter_after_inference = scatter_after_inference
        self.enable_gradient_accumulation = enable_gradient_accumulation
        self.accumulating_grads = False
        self.pin_memory = pin_memory
        self.dev = get_accelerator().get_current_device()
        # ---- chunk configuration
        ws = dist.get_world_size(zero_group) if dist.is_initialized() else 1
        if chunk_config_dict is not None:
            cfg = {k: dict(v) for k, v in chunk_config_dict.items()}
            self.chunk_manager = ChunkManager(cfg, self.dev, max_prefetch=max_prefetch)
        else:
            self.chunk_manager = init_chunk_manager(
                model=module, init_device=self.dev, hidden_dim=hidden_dim, search_range_m=search_range_m,
                min_chunk_size_m=min_chunk_size_m, strict_ddp_flag=strict_ddp_mode, process_group=zero_group,
                verbose=verbose, max_prefetch=max_prefetch)
        self.gemini_manager = GeminiManager(
            placement_policy, self.chunk_manager, memstats, max_prefetch=max_prefetch,
            shard_param_frac=shard_param_frac, offload_optim_frac=offload_optim_frac,
            offload_param_frac=offload_param_frac, warmup_non_model_data_ratio=warmup_non_model_data_ratio,
            steady_cuda_cap_ratio=steady_cuda_cap_ratio)
        self.config_key = list(self.chunk_manager.dp_degree_chunk_size_dict.keys())[0]
        # ---- register parameters: low-precision working chunks + fp32 master chunks with identical layout
        self.param_op_hook_enabled = True
        self.fp16_params: List[nn.Parameter] = []
        self.fp32_params: List[torch.Tensor] = []
        self.name2param: Dict[str, nn.Parameter] = {}
        self.grads_device: Dict[torch.Tensor, torch.device] = {}
        self.overflow_counter = torch.zeros(1, dtype=torch.int32, device=self.dev)
        seen: Set[int] = set()
        offload_master = offload_optim_frac > 0 and get_accelerator().name != "cpu"
        for name, p in module.named_parameters():
            if id(p) in seen or is_ddp_ignored(p):
                continue
            seen.add(id(p))
            self.name2param[name] = p
            fp32 = p.data.detach().float().clone() if master_weights else None
            p.data = p.data.to(self.dev, dtype=mixed_precision)
            self.chunk_manager.register_tensor(p, "fp16_param", self.config_key, zero_group,
                                               extra_dp_group=extra_dp_group, cpu_offload=False, pin_memory=pin_memory)
            self.fp16_params.append(p)
            if master_weights:
                fp32 = fp32.to(self.dev)
                self.chunk_manager.register_tensor(fp32, "fp32_param", self.config_key, zero_group,
                                                   extra_dp_group=extra_dp_group, cpu_offload=False,
                                                   pin_memory=pin_memory)
                self.fp32_params.append(fp32)
        self.chunk_manager.close_all_groups()
        self.param2name = {p: n for n, p in self.name2param.items()}
        self.gemini_manager.setup_grads_device(self.fp16_params, self.grads_device)
        for p, p32 in zip(self.fp16_params, self.fp32_params):
            c16, c32 = self.chunk_manager.get_chunk(p), self.chunk_manager.get_chunk(p32)
            c16.init_pair(c32)
            if self.grads_device[p].type == "cpu" and get_accelerator().name != "cpu":
                self.chunk_manager.move_chunk(c32, torch.device("cpu"))     # optimizer shard lives with its grads
        # static placement of the working chunks
        for c in self._param_chunks():
            if getattr(c, "_static_keep_gathered", False):
                c.keep_gathered = True
                self.chunk_manager.access_chunk(c)
            elif getattr(c, "_static_offload", False):
                self.chunk_manager.move_chunk(c, torch.device("cpu"))
        # buffers follow the module to the accelerator in the working dtype
        for b in module.buffers():
            b.data = b.data.to(self.dev)
            if torch.is_floating_point(b):
                b.data = b.data.to(mixed_precision)
        # ---- hooks on every module that owns parameters directly
        self._handles = []
        self._mod_params: Dict[nn.Module, List[nn.Parameter]] = {}
        for m in module.modules():
            ps = [p for p in m.parameters(recurse=False) if p in self.param2name]
            if not ps:
                continue
            self._mod_params[m] = ps
            self._handles.append(m.register_forward_pre_hook(self._fwd_pre_hook, with_kwargs=True))
            self._handles.append(m.register_forward_hook(self._fwd_post_hook))
        for p in self.fp16_params:
            if p.requires_grad:
                self._handles.append(p.register_post_accumulate_grad_hook(self._grad_handle))
        self._logged = False

    # ------------------------------------------------------------------ helpers
    def _param_chunks(self) -> List[Chunk]:
        return list(dict.fromkeys(self.chunk_manager.get_chunk(p) for p in self.fp16_params))

    def parameters(self, recurse: bool = True):
        return self.module.parameters(recurse)

    def named_parameters(self, prefix: str = "", recurse: bool = True):
        return self.module.named_parameters(prefix, recurse)

    def _access(self, params: List[nn.Parameter]) -> None:
        chunks = self.chunk_manager.get_chunks(params)
        chunks = self.gemini_manager.wait_chunks(chunks) + tuple(c for c in chunks if c in self.chunk_manager.accessed_chunks)
        chunks = tuple(dict.fromkeys(self.chunk_manager.get_chunks(params)))
        self.gemini_manager.sample_overall_data()
        self.gemini_manager.adjust_layout(chunks, record_anyway=self.gemini_manager.placement_policy.max_prefetch > 0)
        for c in chunks:
            self.chunk_manager.access_chunk(c)
        self.gemini_manager.record_model_data_volume()
        # prefetch upcoming chunks
        for c in self.gemini_manager.placement_policy.get_prefetch_chunks(
                self.gemini_manager.is_warmup(), tuple(self.gemini_manager.compute_list),
                self.gemini_manager.compute_idx, self.gemini_manager.async_works):
            self.gemini_manager.add_work(c, self.chunk_manager.access_chunk(c, async_access=True))
        for p in params:
            self.chunk_manager.trans_tensor_state(p, TensorState.COMPUTE)

    def _release(self, params: List[nn.Parameter], state: TensorState) -> None:
        for p in params:
            self.chunk_manager.trans_tensor_state(p, state)
        for c in dict.fromkeys(self.chunk_manager.get_chunks(params)):
            self.chunk_manager.release_chunk(c)

    # ------------------------------------------------------------------ hooks
    def _fwd_pre_hook(self, mod, args, kwargs):
        if not self.param_op_hook_enabled:
            return None
        self._access(self._mod_params[mod])
        if torch.is_grad_enabled():
            needs = [t for t in itertools.chain(args, kwargs.values()) if isinstance(t, torch.Tensor) and t.requires_grad]
            if needs:
                def wrap(t):
                    return _PostBackward.apply(self, mod, t) if (isinstance(t, torch.Tensor) and t.requires_grad) else t
                return _map_tensors(args, wrap), _map_tensors(kwargs, wrap)
        return None

    def _fwd_post_hook(self, mod, args, output):
        if not self.param_op_hook_enabled:
            return None
        self._release(self._mod_params[mod], TensorState.HOLD)
        if torch.is_grad_enabled():
            def wrap(t):
                return _PreBackward.apply(self, mod, t) if (isinstance(t, torch.Tensor) and t.requires_grad) else t
            return _map_tensors(output, wrap)
        return None

    def _pre_backward(self, mod) -> None:
        self._access(self._mod_params[mod])

    def _post_backward_module(self, mod) -> None:
        self._release(self._mod_params[mod], TensorState.HOLD_AFTER_BWD)

    def _grad_handle(self, p: nn.Parameter) -> None:
        grad = p.grad
        if grad is None:
            return
        cm = self.chunk_manager
        chunk = cm.get_chunk(p)
        gchunk = chunk.grad_chunk
        if gchunk is None or (not gchunk.is_gathered and not self.accumulating_grads):
            gchunk = cm.init_grad_chunk(chunk)
        elif not gchunk.is_gathered:
            # gradient accumulation: bring the previously reduced shard back into a gathered buffer
            prev = gchunk.cuda_shard if gchunk.cuda_shard is not None else gchunk.cpu_shard
            gchunk = cm.init_grad_chunk(chunk)
            gchunk.cuda_global_chunk[gchunk.shard_begin:gchunk.shard_end].add_(prev.to(self.dev) * gchunk.pg_size)
        gchunk.add_tensor_to_chunk_slice(p, grad.to(gchunk.dtype))
        p.grad = None
        gchunk.tensor_trans_state(p, TensorState.READY_FOR_REDUCE)
        cm.trans_tensor_state(p, TensorState.HOLD_AFTER_BWD)
        if gchunk.can_reduce:
            cm.reduce_chunk(gchunk)
            flag = gchunk.nonfinite_flag()
            if flag is not None:
                self.overflow_counter += flag.to(self.overflow_counter.device)
            gchunk.set_l2_norm()
            tgt = self.grads_device[p]
            if tgt.type == "cpu" and get_accelerator().name != "cpu":
                cm.move_chunk(gchunk, tgt, force_copy=True)
            cm.release_chunk(chunk)

    # ------------------------------------------------------------------ fwd / bwd
    def forward(self, *args, **kwargs):
        cast = lambda t: t.to(self.mixed_precision) if (isinstance(t, torch.Tensor) and torch.is_floating_point(t)) else t
        args, kwargs = _map_tensors(args, cast), _map_tensors(kwargs, cast)
        self.module.zero_grad(set_to_none=True)
        self.gemini_manager.pre_iter()
        outputs = self.module(*args, **kwargs)
        if not torch.is_grad_enabled() and self.scatter_after_inference:
            self._post_forward()
        if self.force_outputs_fp32:
            outputs = _map_tensors(outputs, lambda t: t.float() if torch.is_floating_point(t) else t)
        return outputs

    def _post_forward(self) -> None:
        for c in self._param_chunks():
            for t in c.get_tensors():
                c.tensor_trans_state(t, TensorState.HOLD)
            self.chunk_manager.release_chunk(c)
        self.gemini_manager.post_iter()

    def _post_backward(self) -> None:
        for c in self._param_chunks():
            for t in c.get_tensors():
                if c.tensors_info[t].state in (TensorState.COMPUTE, TensorState.HOLD_AFTER_BWD):
                    c.tensor_trans_state(t, TensorState.HOLD_AFTER_BWD if c.tensors_info[t].state == TensorState.COMPUTE
                                         else TensorState.HOLD_AFTER_BWD)
            for t in c.get_tensors():
                if c.tensors_info[t].state == TensorState.HOLD_AFTER_BWD:
                    c.tensors_info[t].state = TensorState.HOLD
            c.tensor_state_cnter = {s: 0 for s in TensorState}
            c.tensor_state_cnter[TensorState.HOLD] = c.num_tensors
            self.chunk_manager.release_chunk(c)
        if not self._logged and self.gemini_manager.policy_name == "auto":
            self._logged = True
        if self.enable_gradient_accumulation:
            # from now on backward passes add into the reduced shards until the optimizer steps (reference
            # gemini_ddp.py:341-342 turns the state on here, gemini_optimizer.py:293 turns it off after the step)
            self.accumulating_grads = True
        self.gemini_manager.post_iter()

    def backward(self, loss: torch.Tensor) -> None:
        loss.backward()
        self._post_backward()

    def backward_by_grad(self, tensor, grad, inputs=None, retain_graph=False) -> None:
        torch.autograd.backward(tensor, grad, inputs=inputs, retain_graph=retain_graph)
        self._post_backward()

    def set_chunk_grad_device(self, chunk: Chunk, device: torch.device) -> None:
        for t in chunk.get_tensors():
            self.grads_device[t] = device

    # ------------------------------------------------------------------ state dict
    def _gather_chunk_tensors(self, chunk: Chunk, dtype: Optional[torch.dtype] = None) -> Dict[torch.Tensor, torch.Tensor]:
        """{member tensor: full value} for one chunk (temporarily gathers it)."""
        was = chunk.is_gathered
        if not was:
            dev = self.dev
            shard = chunk.cuda_shard if chunk.cuda_shard is not None else chunk.cpu_shard.to(dev)
            full = torch.empty(chunk.chunk_size, dtype=chunk.dtype, device=dev)
            if chunk.pg_size > 1:
                dist.all_gather_into_tensor(full, shard.contiguous(), group=chunk.torch_pg)
            else:
                full.copy_(shard)
        else:
            full = chunk.cuda_global_chunk
        out = {}
        for t, info in chunk.tensors_info.items():
            v = full[info.offset:info.end].view(info.shape).clone()
            out[t] = v.to(dtype) if dtype is not None else v
        return out

    def state_dict(self, destination=None, prefix: str = "", keep_vars: bool = False, only_rank_0: bool = True,
                   dtype: torch.dtype = None):
        dst = OrderedDict() if destination is None else destination
        dtype = dtype or self.mixed_precision
        use_master = self.master_weights and len(self.fp32_params) == len(self.fp16_params)
        p2master = dict(zip(self.fp16_params, self.fp32_params)) if use_master else {}
        cache: Dict[Chunk, Dict] = {}
        rank0 = (not dist.is_initialized()) or dist.get_rank(self.zero_group) == 0
        for name, p in self.module.named_parameters():
            if p not in self.param2name:
                dst[prefix + name] = p if keep_vars else p.detach()
                continue
            src = p2master.get(p, p)
            chunk = self.chunk_manager.get_chunk(src)
            if chunk not in cache:
                cache.clear()
                cache[chunk] = self._gather_chunk_tensors(chunk, dtype)
            if rank0 or not only_rank_0:
                dst[prefix + name] = cache[chunk][src].cpu()
        # a parameter registered under several names (tied head, a module reused twice) appears under every one of them
        # in `nn.Module.state_dict`; `named_parameters()` only yields the first - add the aliases (same tensor)
        first = {id(p): n for n, p in self.module.named_parameters()}
        for name, p in self.module.named_parameters(remove_duplicate=False):
            canon = first[id(p)]
            if name != canon and prefix + canon in dst and prefix + name not in dst:
                dst[prefix + name] = dst[prefix + canon]
        for name, b in self.module.named_buffers():
            mod_path, _, bname = name.rpartition(".")
            owner = self.module.get_submodule(mod_path) if mod_path else self.module
            if bname not in owner._non_persistent_buffers_set:
                dst[prefix + name] = b if keep_vars else b.detach()
        return dst

    def load_state_dict(self, state_dict: "OrderedDict[str, torch.Tensor]", strict: bool = True):
        known = dict(self.module.named_parameters(remove_duplicate=False))
        missing, unexpected = [], [k for k in state_dict if k not in known and k not in dict(self.module.named_buffers())]
        p2master = dict(zip(self.fp16_params, self.fp32_params)) if self.fp32_params else {}
        for name, p in self.module.named_parameters():
            if name not in state_dict:
                missing.append(name)
                continue
            v = state_dict[name]
            if p not in self.param2name:
                with torch.no_grad():
                    p.copy_(v)
                continue
            for tensor in [t for t in (p, p2master.get(p)) if t is not None]:
                chunk = self.chunk_manager.get_chunk(tensor)
                info = chunk.tensors_info[tensor]
                flat = v.reshape(-1).to(chunk.dtype)
                if chunk.is_gathered:
                    chunk.cuda_global_chunk[info.offset:info.end].copy_(flat)
                else:
                    s, e = max(info.offset, chunk.shard_begin), min(info.end, chunk.shard_end)
                    if e > s:
                        shard = chunk.cuda_shard if chunk.cuda_shard is not None else chunk.cpu_shard
                        shard[s - chunk.shard_begin:e - chunk.shard_begin].copy_(flat[s - info.offset:e - info.offset])
        for name, b in self.module.named_buffers():
            if name in state_dict:
                with torch.no_grad():
                    b.copy_(state_dict[name])
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing}, unexpected {unexpected}")
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def state_dict_shard(self, prefix: str = "", keep_vars: bool = False, max_shard_size: int = 1024,
                         only_rank_0: bool = True, dtype: torch.dtype = torch.float16, pinned_state_dicts=None
                         ) -> Iterator[Tuple[OrderedDict, int]]:
        from ...checkpoint_io.utils import StateDictSharder

        sharder = StateDictSharder(max_shard_size)
        seen = set()
        for k, v in self.state_dict(prefix=prefix, only_rank_0=only_rank_0, dtype=dtype).items():
            if torch.is_tensor(v):
                if id(v) in seen:               # alias of a tied / shared parameter: stored once (HF convention)
                    continue
                seen.add(id(v))
            block, size = sharder.append_param(k, v)
            if block is not None:
                yield block, size
        yield sharder.current_block, sharder.current_block_size
