"""`fp8_communication=True` threaded through the tensor / sequence parallel layers: collectives carry fp8 payloads
(forward e4m3 activations, backward e5m2 gradients), results stay within fp8 error of the bf16 run (reference:
tests/test_fp8/test_fp8_{allgather,allreduce,reduce_scatter}.py + the `fp8_communication` plugin flag)."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.cluster import DeviceMesh
from colossalai_b200.models import build_model
from colossalai_b200.parallel import comm
from colossalai_b200.shardformer import ShardConfig, ShardFormer
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    mesh = DeviceMesh(dp=1, tp=2)
    g = mesh.group("tp")
    # the comm wrappers honour the switch
    x = torch.randn(4, 64, generator=torch.Generator().manual_seed(rank))
    exact = comm.all_gather(x, 0, g)
    with comm.fp8_communication(True):
        lossy = comm.all_gather(x, 0, g)
        rs = comm.reduce_scatter(exact.clone(), 0, g)
        ar = comm.all_reduce(x.clone(), g)
    assert not torch.equal(lossy, exact) and (lossy - exact).abs().max() < 0.13 * exact.abs().max()
    torch.testing.assert_close(rs, 2 * exact[rank * 4:(rank + 1) * 4], atol=0.5, rtol=0.15)
    torch.testing.assert_close(ar, exact[:4] + exact[4:], atol=0.4, rtol=0.1)
    assert not comm.fp8_enabled()
    for sp_mode in (None, "split_gather"):
        torch.manual_seed(1234)
        base = build_model("llama-tiny")
        ids = torch.randint(0, 512, (2, 32), generator=torch.Generator().manual_seed(7))
        outs = {}
        for fp8 in (False, True):
            sc = ShardConfig(tensor_parallel_process_group=g, enable_tensor_parallelism=True,
                             enable_sequence_parallelism=sp_mode is not None, sequence_parallelism_mode=sp_mode,
                             fp8_communication=fp8)
            m, _ = ShardFormer(sc).optimize(copy.deepcopy(base))
            loss = m(input_ids=ids, labels=ids)["loss"]
            loss.backward()
            gn = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in m.parameters() if p.grad is not None))
            outs[fp8] = (loss.item(), gn.item())
        (l0, g0), (l1, g1) = outs[False], outs[True]
        assert l0 != l1, "fp8 communication must actually change the payloads"
        assert abs(l1 - l0) < 0.05 * abs(l0) and abs(g1 - g0) < 0.25 * g0, (sp_mode, outs)
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_fp8_communication_in_tp_sp_layers():
    spawn(_worker, 2)
