"""2D / 2.5D / 3D tensor-parallel linears == dense linear (values and gradients) on gloo meshes
(reference: tests/test_legacy/test_layers/test_{2d,2p5d,3d})."""
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.legacy import ParallelMode, global_context as gpc
from colossalai_b200.legacy.nn.layer import Linear2D, Linear2p5D, Linear3D, split_2d, split_2p5d, split_3d_input
from colossalai_b200.legacy.nn.layer.parallel_3d import gather_3d_output, split_3d_weight
from colossalai_b200.parallel import comm
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn

M, K, N = 16, 8, 12


def _dense():
    torch.manual_seed(0)
    x = torch.randn(M, K)
    w = torch.randn(K, N)
    b = torch.randn(N)
    dy = torch.randn(M, N)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    y = xr @ wr + br
    y.backward(dy)
    return x, w, b, dy, y.detach(), xr.grad, wr.grad, br.grad


def _w2d(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    gpc.init_tensor_mesh("2d", 4)
    x, w, b, dy, y, gx, gw, gb = _dense()
    layer = Linear2D(K, N, bias=True)
    j = gpc.get_local_rank(ParallelMode.PARALLEL_2D_ROW)
    with torch.no_grad():
        layer.weight.copy_(split_2d(w))
        layer.bias.copy_(b.chunk(2)[j])
    xl = split_2d(x).requires_grad_()
    out = layer(xl)
    torch.testing.assert_close(out, split_2d(y), rtol=1e-4, atol=1e-5)
    out.backward(split_2d(dy))
    torch.testing.assert_close(xl.grad, split_2d(gx), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.weight.grad, split_2d(gw), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.bias.grad, gb.chunk(2)[j], rtol=1e-4, atol=1e-5)
    dist.barrier()
    dist.destroy_process_group()


def _w2p5d(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    gpc.init_tensor_mesh("2.5d", 8, depth=2)
    x, w, b, dy, y, gx, gw, gb = _dense()
    layer = Linear2p5D(K, N, bias=True)
    j = gpc.get_local_rank(ParallelMode.PARALLEL_2P5D_ROW)
    with torch.no_grad():
        layer.weight.copy_(split_2d_like(w))
        layer.bias.copy_(b.chunk(2)[j])
    xl = split_2p5d(x).requires_grad_()
    out = layer(xl)
    torch.testing.assert_close(out, split_2p5d(y), rtol=1e-4, atol=1e-5)
    out.backward(split_2p5d(dy))
    torch.testing.assert_close(xl.grad, split_2p5d(gx), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.weight.grad, split_2d_like(gw), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.bias.grad, gb.chunk(2)[j], rtol=1e-4, atol=1e-5)
    dist.barrier()
    dist.destroy_process_group()


def split_2d_like(w):
    """weight block [i, j] of the 2.5D mesh (identical on every depth layer)"""
    i = gpc.get_local_rank(ParallelMode.PARALLEL_2P5D_COL)
    j = gpc.get_local_rank(ParallelMode.PARALLEL_2P5D_ROW)
    return w.chunk(2, 0)[i].chunk(2, 1)[j].contiguous()


def _w3d(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    gpc.init_tensor_mesh("3d", 8)
    x, w, b, dy, y, gx, gw, gb = _dense()
    y = y - b                      # Linear3D carries no bias
    layer = Linear3D(K, N)
    with torch.no_grad():
        layer.weight.copy_(split_3d_weight(w))
    xl = split_3d_input(x).requires_grad_()
    out = layer(xl)
    torch.testing.assert_close(gather_3d_output(out), y, rtol=1e-4, atol=1e-5)
    # output layout: rows (i then j), cols k
    i = gpc.get_local_rank(ParallelMode.PARALLEL_3D_WEIGHT)
    j = gpc.get_local_rank(ParallelMode.PARALLEL_3D_INPUT)
    k = gpc.get_local_rank(ParallelMode.PARALLEL_3D_OUTPUT)
    out.backward(dy.chunk(2, 0)[i].chunk(2, 0)[j].chunk(2, 1)[k].contiguous())
    torch.testing.assert_close(xl.grad, split_3d_input(gx), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.weight.grad, split_3d_weight(gw), rtol=1e-4, atol=1e-5)
    dist.barrier()
    dist.destroy_process_group()


@rerun_if_address_is_in_use()
def test_linear_2d():
    spawn(_w2d, 4)


@rerun_if_address_is_in_use()
def test_linear_2p5d():
    spawn(_w2p5d, 8)


@rerun_if_address_is_in_use()
def test_linear_3d():
    spawn(_w3d, 8)
