"""CPU (gloo, world_size 2): the data-parallel host logic -- batch sharding, global-batch loss scaling and the single
mesh-head gradient all-reduce.  The warp itself needs a GPU; here the per-rank warp is played by the
oracle port so that what is tested is the plumbing: N-rank result == 1-rank result on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dovs_b200
from dovs_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, out_q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'oracle'))
    import mesh_warp_ref as ref
    import synth
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    torch.set_num_threads(1)
    h, wd, c, gh, gw, nf = 24, 32, 1, 4, 4, 16
    U = torch.tensor(synth.smooth_image(n_global, h, wd, c, 1))
    y = torch.tensor(synth.smooth_image(n_global, h, wd, c, 2))
    theta = torch.tensor(synth.random_mesh(n_global, gh, gw, 0.03, 3))
    feats = torch.tensor(synth.randn((n_global, nf), 4))
    Ul, yl, fl = (parallel.shard(t, rank, world) for t in (U, y, feats))
    thl = parallel.shard(theta, rank, world).requires_grad_(True)
    out, black, img, _ = ref.transformer(Ul, thl)
    nb = 1 - black.reshape(-1, h, wd, 1)
    err = (out - yl) * nb
    # per-sample sums divided by the GLOBAL batch (reference s_net_bundle_nobm.py:352 divides by batch_size)
    loss = ((err * err).sum((1, 2, 3)) / (nb.sum((1, 2, 3)) + 1e-8)).sum() / n_global
    loss.backward()
    red = parallel.MeshHeadGradReducer(nf, 2 * (gh + 1) * (gw + 1), 'cpu')
    red.head_grad(fl, thl.grad)
    red.launch()
    buf = red.wait().clone()
    lo, hi = parallel.shard_bounds(n_global, rank, world)
    out_q.put((rank, lo, hi, buf.numpy(), float(loss)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_global', [6, 5])
def test_two_ranks_equal_one_rank(n_global):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    q1 = ctx.Queue()
    p1 = ctx.Process(target=_worker, args=(0, 1, _free_port(), n_global, q1))
    p1.start()
    single = q1.get(timeout=180)
    p1.join(timeout=60)
    # shards partition the batch, every rank holds the same reduced gradient, and it equals the 1-rank gradient
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_global
    assert np.array_equal(res[0][3], res[1][3])
    assert np.abs(res[0][3] - single[3]).max() <= 1e-5 * max(1.0, np.abs(single[3]).max())
    assert abs(res[0][4] + res[1][4] - single[4]) <= 1e-6 * max(1.0, abs(single[4]))


def test_shard_bounds_cover_everything():
    for n in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def _grad_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_from_env(backend='gloo')
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    x = torch.arange(6 * 5, dtype=torch.float32).reshape(6, 5) / 10
    xl = parallel.shard(x, rank, world)
    (lin(xl).square().sum() / x.shape[0]).backward()                 # divided by the GLOBAL batch
    parallel.allreduce_grads(lin.parameters(), bucket_bytes=64)      # several buckets
    out_q.put((rank, [p.grad.numpy().copy() for p in lin.parameters()]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_allreduce_grads_two_ranks_equal_one_rank():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    q1 = ctx.Queue()
    p1 = ctx.Process(target=_grad_worker, args=(0, 1, _free_port(), q1))
    p1.start()
    single = q1.get(timeout=180)
    p1.join(timeout=60)
    for a, b, c in zip(res[0][1], res[1][1], single[1]):
        assert np.array_equal(a, b) and np.allclose(a, c, rtol=1e-5, atol=1e-6)


def test_stabnet_carrier_shapes_and_head_init():
    """the backbone + head carrier (s_net_bundle_nobm.py:250-264) on CPU: theta shape, output_layer initialiser (resnet.py:50-53)"""
    torch.manual_seed(0)
    net = dovs_b200.StabNet(in_ch=13, grid=(4, 4)).eval()
    with torch.no_grad():
        f = net.backbone(torch.zeros(1, 13, 64, 96))
        theta = net(torch.zeros(2, 64, 96, 13))
    assert f.shape == (1, 2048, 2, 3)                                # output_stride 32
    assert theta.shape == (2, 50)
    assert float(net.head.weight.abs().max()) <= (3.0 / 512) ** 0.5 and float(net.head.bias.abs().max()) == 0.0
    n_units = sum(1 for m in net.backbone.units)
    assert n_units == 16
    assert dovs_b200.StabNet(in_ch=13, grid=(2, 3)).head.out_features == 24
