"""Legacy pipeline: PipelinableContext partitioning and the RPC engines (reference: tests/test_legacy/test_pipeline/
{test_pipelinable.py, rpc_test_utils.py, test_cuda_rpc_pipeline.py})."""
import os

import pytest
import torch
import torch.nn as nn

from colossalai_b200.legacy.pipeline import (PipelinableContext, partition_balanced, partition_uniform)
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


class MLP(nn.Module):
    def __init__(self, dim=16, layers=6):
        super().__init__()
        for i in range(layers):
            setattr(self, f"fc{i}", nn.Linear(dim, dim))

    def forward(self, x):
        for _, m in self.named_children():
            x = torch.tanh(m(x))
        return x


def test_partition_helpers_and_pipelinable():
    assert partition_uniform(10, 4) == [[(0, 3)], [(3, 6)], [(6, 8)], [(8, 10)]]
    parts = partition_balanced([1, 1, 1, 1, 8, 1, 1, 1], 3)
    spans = sorted(p[0] for p in parts)
    assert spans[0][0] == 0 and spans[-1][1] == 8 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    w = [1, 1, 1, 1, 8, 1, 1, 1]
    assert max(sum(w[s:e]) for s, e in spans) <= 9
    torch.manual_seed(0)
    ctx = PipelinableContext(policy="uniform")
    with ctx:
        model = MLP()
    ctx.to_layer_list([f"fc{i}" if j == 0 else torch.tanh for i in range(6) for j in range(2)])
    assert ctx.layers_count == 6
    x = torch.randn(3, 16)
    stages = [ctx.partition(1, 2, r) for r in range(2)]
    torch.testing.assert_close(stages[1](stages[0](x)), model(x))


def _stage_fn(stage: int) -> nn.Module:
    torch.manual_seed(100 + stage)
    return nn.Sequential(nn.Linear(8, 8), nn.Tanh(), nn.Linear(8, 8 if stage == 0 else 2))


def _rpc_worker(rank, world_size, port):
    import torch.distributed.rpc as rpc

    from colossalai_b200.legacy.pipeline import FillDrainPipelineEngine, OneFOneBPipelineEngine

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    opts = rpc.TensorPipeRpcBackendOptions(num_worker_threads=8, init_method=f"tcp://127.0.0.1:{port}", rpc_timeout=60)
    rpc.init_rpc(f"work{rank}", rank=rank, world_size=world_size, rpc_backend_options=opts)
    if rank == 0:
        x = torch.randn(8, 8, generator=torch.Generator().manual_seed(1))
        y = torch.randn(8, 2, generator=torch.Generator().manual_seed(2))
        ref = nn.Sequential(_stage_fn(0), _stage_fn(1))
        loss_ref = sum(nn.functional.mse_loss(ref(xc), yc) for xc, yc in zip(x.chunk(4), y.chunk(4)))
        loss_ref.backward()
        for cls in (FillDrainPipelineEngine, OneFOneBPipelineEngine):
            eng = cls(_stage_fn, stage_num=2, num_microbatches=4, criterion=nn.functional.mse_loss)
            losses = eng.forward_backward(x, y)
            torch.testing.assert_close(sum(losses), loss_ref.detach(), atol=1e-5, rtol=1e-5)
            grads = eng.remote_parameters_grads()
            for s in range(2):
                for n, p in ref[s].named_parameters():
                    torch.testing.assert_close(grads[s][n], p.grad, atol=1e-5, rtol=1e-4)
            eng.initialize_optimizer(torch.optim.SGD, lr=0.1)
            eng.step()
            out = eng.forward_backward(x, forward_only=True)
            assert len(out) == 4 and out[0].shape == (2, 2)
    rpc.shutdown()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_rpc_pipeline_engines_match_single_process():
    spawn(_rpc_worker, 2)
