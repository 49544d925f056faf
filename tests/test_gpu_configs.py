"""GPU parity at the BASELINE.json configurations that the fixture-driven tests do not reach, and the production
(multi-iteration / C = 3) code paths of the loss kernels.  All checks are against the oracles (oracle/mgw_oracle.c
bit-exact forward; oracle/mesh_warp_ref.py fp64 autograd), on seeded synthetic inputs, through the public API /
the C ABI.  Reference expressions: spatial_transformer3.py:19-365, s_net_bundle_nobm.py:335-352,
train_bundle_nobm.py:115-125.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import c_oracle
import mesh_warp_ref as ref
import synth
from conftest import ROOT, bits_equal, relmax

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mgw():
    import dovs_b200
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    yield dovs_b200
    dovs_b200.set_impl('auto')


def dev(a, dtype=torch.float32):
    return torch.tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


def t64(a, grad=False):
    return torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=grad)


def close_but_knife_edges(got, ref64, tol, max_outliers):
    """max-norm-relative agreement with the fp64 port EXCEPT on a handful of knife-edge elements.  The reference function is
    discontinuous where a sample point crosses the image border (weights come from the CLIPPED integers,
    spatial_transformer3.py:90-93,114-121: the four terms cancel outside), so a pixel whose coordinate lies within fp32
    rounding of W-1 / H-1 contributes its whole gradient in one precision and nothing in the other -- the reference's own
    fp32 and fp64 runs differ on 9 of 2 M pixels at 1080p (by up to 24 % of max|dU|).  Returns (ok, #outliers, worst inlier)."""
    got, ref64 = np.asarray(got, np.float64), np.asarray(ref64, np.float64)
    d = np.abs(got - ref64) / max(np.abs(ref64).max(), 1e-30)
    bad = d > tol
    return int(bad.sum()) <= max_outliers, int(bad.sum()), float(d[~bad].max()) if (~bad).any() else 0.0


# ------------------------------------------------------------------ config #4: 1 x 1080 x 1920 x 3 (cells 270 x 480: ragged tile rows)
@pytest.mark.parametrize('impl', ['auto', 'generic', 'tma'])
@pytest.mark.parametrize('sigma', [0.0, 0.03])
def test_config4_1080p_forward_bit_exact(mgw, impl, sigma):
    mgw.set_impl(impl)
    n, h, w, c = 1, 1080, 1920, 3
    U = synth.noise_image(n, h, w, c, 410)
    theta = synth.random_mesh(n, 4, 4, sigma, 411)
    out, black, img, Hs = mgw.ops.mesh_warp_fwd(dev(U), dev(theta))
    o_Hs = c_oracle.solve_h(theta)
    assert bits_equal(Hs.cpu().numpy(), o_Hs).all()
    o_out, o_black, o_img, _ = c_oracle.warp(U, o_Hs)
    assert bits_equal(out.cpu().numpy(), o_out).all()
    assert (black.cpu().numpy() == o_black).all()
    assert bits_equal(img.cpu().numpy(), o_img).all()


@pytest.mark.parametrize('impl', ['auto', 'generic', 'tma'])
def test_config4_1080p_backward(mgw, impl):
    """dU / dHs at 1080p against the fp64 autograd of the oracle port at the SAME Hs (smooth image: the gradient is
    continuous in the sample position there)."""
    mgw.set_impl(impl)
    n, h, w, c = 1, 1080, 1920, 3
    U = synth.smooth_image(n, h, w, c, 412)
    theta = synth.random_mesh(n, 4, 4, 0.03, 413)
    g, gi = synth.randn((n, h, w, c), 414), synth.randn((n, h, w, 2), 415, 0.1)
    Hs = mgw.ops.solve_h_fwd(dev(theta))
    dU, dHs = mgw.ops.warp_bwd(dev(U), Hs, dev(g), dev(gi))
    # the reference's own precision is the yardstick at this size: x reaches 1920, where one fp32 ulp of the sample position is
    # 1.2e-4 px, i.e. ~1e-4 of max|dU| per tap weight -- the fp32 port (same operation order) against its own fp64 run is no closer
    U64, H64 = t64(U, True), Hs.cpu().double().requires_grad_(True)
    o, _, im = ref.warp_given_h(U64, H64)
    torch.autograd.backward([o, im], [t64(g), t64(gi)])
    U32, H32 = torch.tensor(U, requires_grad=True), Hs.cpu().clone().requires_grad_(True)
    o, _, im = ref.warp_given_h(U32, H32)
    torch.autograd.backward([o, im], [torch.tensor(g), torch.tensor(gi)])
    ok, nbad, worst = close_but_knife_edges(dU.cpu().numpy(), U32.grad.numpy(), 1e-4, 64 * c)
    assert ok, (nbad, worst)
    d64 = np.abs(dU.cpu().numpy().astype(np.float64) - U64.grad.numpy()) / np.abs(U64.grad.numpy()).max()
    r64 = np.abs(U32.grad.numpy().astype(np.float64) - U64.grad.numpy()) / np.abs(U64.grad.numpy()).max()
    assert np.quantile(d64, 0.9999) <= max(1e-4, 2 * np.quantile(r64, 0.9999)), (np.quantile(d64, 0.9999), np.quantile(r64, 0.9999))
    # dHs sums every pixel of a cell, the knife-edge ones included
    h64, h32, got = H64.grad.numpy()[..., :8], H32.grad.numpy()[..., :8], dHs.cpu().numpy()[..., :8]
    assert relmax(got, h64) <= max(1e-4, 2 * relmax(h32, h64)), (relmax(got, h64), relmax(h32, h64))
    assert relmax(got, h32) <= max(1e-4, 2 * relmax(h32, h64))


# ------------------------------------------------------------------ config #3: 16 x 288 x 512 x 1 (the network's own call, s_net_bundle_nobm.py:281,332)
@pytest.mark.parametrize('impl', ['auto', 'generic', 'tma'])
def test_config3_c1_forward_and_backward(mgw, impl):
    mgw.set_impl(impl)
    n, h, w, c = 16, 288, 512, 1
    U = synth.noise_image(n, h, w, c, 310)
    theta = synth.random_mesh(n, 4, 4, 0.05, 311)
    out, black, img, Hs = mgw.ops.mesh_warp_fwd(dev(U), dev(theta))
    o_Hs = c_oracle.solve_h(theta)
    assert bits_equal(Hs.cpu().numpy(), o_Hs).all()
    for k in (0, 7, 15):
        o_out, o_black, o_img, _ = c_oracle.warp(U[k:k + 1], o_Hs[k:k + 1])
        assert bits_equal(out[k:k + 1].cpu().numpy(), o_out).all()
        assert (black[k:k + 1].cpu().numpy() == o_black).all() and bits_equal(img[k:k + 1].cpu().numpy(), o_img).all()
    # backward on the whole batch, checked on a 2-sample slice against fp64 autograd at the same Hs
    Us = synth.smooth_image(n, h, w, c, 312)
    g, gi = synth.randn((n, h, w, c), 313), synth.randn((n, h, w, 2), 314, 0.1)
    dU, dth = mgw.ops.mesh_warp_bwd(dev(Us), dev(theta), Hs, dev(g), dev(gi))
    dU2, dHs = mgw.ops.warp_bwd(dev(Us), Hs, dev(g), dev(gi))
    assert torch.equal(dU, dU2) or relmax(dU.cpu().numpy(), dU2.cpu().numpy()) < 1e-6
    sl = slice(3, 5)
    U64, H64 = t64(Us[sl], True), Hs[sl].cpu().double().requires_grad_(True)
    o, _, im = ref.warp_given_h(U64, H64)
    torch.autograd.backward([o, im], [t64(g[sl]), t64(gi[sl])])
    assert relmax(dU[sl].cpu().numpy(), U64.grad.numpy()) < 1e-4
    assert relmax(dHs[sl].cpu().numpy()[..., :8], H64.grad.numpy()[..., :8]) < 1e-4
    th64 = t64(theta[sl], True)
    up = dHs[sl].cpu().double()
    up[..., 8] = 0
    (ref.solve_h(th64) * up).sum().backward()
    # the adjoint solve on the SAME dHs (K4 alone), then the fused call's dtheta against it: the two calls are separate runs, and
    # the generic family sums dHs with fp32 atomics in arbitrary order -- noise the adjoint solve amplifies by the conditioning
    # of the cell's 8x8 system
    dth2 = mgw.ops.solve_h_bwd(dev(theta), Hs, dHs)
    assert relmax(dth2[sl].cpu().numpy(), th64.grad.numpy()) < 1e-4
    assert relmax(dth.cpu().numpy(), dth2.cpu().numpy()) < (2e-3 if impl == 'generic' else 1e-4)


# ------------------------------------------------------------------ loss epilogues on their PRODUCTION code paths
# 4 x 288 x 512: the grid-stride loops of mgw_loss.cu run several iterations per thread (every fixture runs one), C = 3
# instantiates the templates no fixture reaches, 3000 matches is the reference's own count (s_net_bundle_nobm.py:277-278).
def _loss_inputs(n, h, w, c, seed):
    r = np.random.RandomState(seed)
    out1, out2 = synth.smooth_image(n, h, w, c, seed + 1), synth.smooth_image(n, h, w, c, seed + 2)
    y = synth.smooth_image(n, h, w, c, seed + 3)
    black1 = (r.rand(n, h, w) < 0.1).astype(np.float32)
    black2 = (r.rand(n, h, w) < 0.1).astype(np.float32)
    # a smooth optical-flow-like field in normalised coordinates with some samples leaving the frame
    yy, xx = np.meshgrid(np.linspace(-1, 1, h, dtype=np.float32), np.linspace(-1, 1, w, dtype=np.float32), indexing='ij')
    flow = np.stack([xx[None] + 0.08 * np.sin(3 * yy + np.arange(n)[:, None, None]) + 0.01 * r.randn(n, h, w),
                     yy[None] * 1.03 + 0.05 * np.cos(2 * xx + np.arange(n)[:, None, None])], axis=-1).astype(np.float32)
    return out1, out2, y, black1, black2, flow


@pytest.mark.parametrize('c', [1, 3])
def test_img_loss_production_path(mgw, c):
    n, h, w = 4, 288, 512
    out1, _, y, black1, _, _ = _loss_inputs(n, h, w, c, 500 + c)
    o, yy, b = dev(out1).requires_grad_(True), dev(y), dev(black1)
    loss = mgw.img_loss(o, yy, b)
    loss.backward()
    o64 = t64(out1, True)
    l64 = ref.img_loss(o64, t64(y), t64(black1))
    l64.backward()
    assert abs(float(loss) - float(l64)) <= 1e-5 * abs(float(l64))
    assert relmax(o.grad.cpu().numpy(), o64.grad.numpy()) < 1e-5


@pytest.mark.parametrize('c', [1, 3, 4])
def test_temp_loss_production_path(mgw, c):
    n, h, w = 4, 288, 512
    out1, out2, _, black1, black2, flow = _loss_inputs(n, h, w, c, 520 + c)
    a, b2 = dev(out1).requires_grad_(True), dev(out2).requires_grad_(True)
    loss = mgw.temp_loss(a, dev(black1), b2, dev(black2), dev(flow))
    loss.backward()
    a64, b64 = t64(out1, True), t64(out2, True)
    l64 = ref.temp_loss(a64, t64(black1), b64, t64(black2), t64(flow))
    l64.backward()
    assert abs(float(loss) - float(l64)) <= 2e-5 * abs(float(l64))
    assert relmax(a.grad.cpu().numpy(), a64.grad.numpy()) < 1e-4
    assert relmax(b2.grad.cpu().numpy(), b64.grad.numpy()) < 1e-4


@pytest.mark.parametrize('c', [1, 3])
@pytest.mark.parametrize('wild', [False, True])
def test_temp_loss_partial_tiles_and_fallback(mgw, c, wild):
    """the tiled temp_loss on an image that is no multiple of its 32 x 24 tiles, with a compact flow (boxes fit: shared-memory
    path) and with a flow that jumps all over the frame (boxes do not fit: per-pixel fallback inside the same kernel), forward
    and backward against the fp64 port"""
    n, h, w = 2, 100, 132
    out1, out2, _, black1, black2, flow = _loss_inputs(n, h, w, c, 620 + c)
    if wild:
        flow = np.random.RandomState(7).uniform(-1.2, 1.2, flow.shape).astype(np.float32)
    a, b2 = dev(out1).requires_grad_(True), dev(out2).requires_grad_(True)
    loss = mgw.temp_loss(a, dev(black1), b2, dev(black2), dev(flow))
    loss.backward()
    a64, b64 = t64(out1, True), t64(out2, True)
    l64 = ref.temp_loss(a64, t64(black1), b64, t64(black2), t64(flow))
    l64.backward()
    assert abs(float(loss) - float(l64)) <= 2e-5 * abs(float(l64))
    assert relmax(a.grad.cpu().numpy(), a64.grad.numpy()) < 1e-4
    assert relmax(b2.grad.cpu().numpy(), b64.grad.numpy()) < 1e-4


def test_feature_loss_production_path(mgw):
    n, h, w, m = 4, 288, 512, 3000
    r = np.random.RandomState(540)
    img = synth.randn((n, h, w, 2), 541, 0.5)
    matches = r.uniform(-1.1, 1.1, (n, m, 4)).astype(np.float32)          # some points clip to the border (:216-221)
    mask = (r.rand(n, m) < 0.3).astype(np.float32)
    mask[3] = 0                                                            # a sample without matches: divisor max(sum,1)
    f = dev(img).requires_grad_(True)
    loss, warpped = mgw.feature_loss(dev(matches), dev(mask), f)
    loss.backward()
    f64 = t64(img, True)
    l64, w64 = ref.feature_loss(t64(matches), t64(mask), f64)
    l64.backward()
    assert np.array_equal(warpped.cpu().numpy(), w64.detach().numpy().astype(np.float32))
    assert abs(float(loss) - float(l64)) <= 1e-5 * abs(float(l64))
    # coincident match points accumulate in fp32 atomics: order-dependent rounding only
    assert relmax(f.grad.cpu().numpy(), f64.grad.numpy()) < 1e-5


@pytest.mark.parametrize('c', [1, 3])
def test_fused_img_loss_production_path(mgw, c):
    """transformer + img_loss fused into the warp kernels at 4 x 288 x 512 against the fp64 port (value and dtheta / dU)."""
    n, h, w = 4, 288, 512
    U, y = synth.smooth_image(n, h, w, c, 560 + c), synth.smooth_image(n, h, w, c, 570 + c)
    theta = synth.random_mesh(n, 4, 4, 0.05, 580 + c)
    Ua, ta = dev(U).requires_grad_(True), dev(theta).requires_grad_(True)
    loss, out, black, img = mgw.transformer_img_loss(Ua, ta, dev(y))
    loss.backward()
    # the same stages at the SAME Hs in fp64
    Hs = mgw.ops.solve_h_fwd(dev(theta))
    U64, H64 = t64(U, True), Hs.cpu().double().requires_grad_(True)
    o64, b64, _ = ref.warp_given_h(U64, H64)
    l64 = ref.img_loss(o64, t64(y), b64)
    l64.backward()
    assert abs(float(loss) - float(l64)) <= 2e-5 * abs(float(l64))
    assert relmax(Ua.grad.cpu().numpy(), U64.grad.numpy()) < 1e-4
    th64 = t64(theta, True)
    up = H64.grad.clone()
    up[..., 8] = 0
    (ref.solve_h(th64) * up).sum().backward()
    assert relmax(ta.grad.cpu().numpy(), th64.grad.numpy()) < 2e-4


@pytest.mark.parametrize('n', [1, 7, 4096, 1080 * 1920 * 3 + 3])
def test_frame_transport_u8(mgw, n):
    """uint8 <-> the network's fp32 range on the device (config.py:19, deploy_bundle.py:75): exact / byte-exact vs the literal
    numpy expressions, odd lengths and unaligned views included."""
    import deploy_ref
    r = np.random.RandomState(n % 1000)
    frame = r.randint(0, 256, n + 1).astype(np.uint8)
    for off in (0, 1):                                         # off = 1: a view whose base is not 4-byte aligned
        src = frame[off:off + n]
        got = mgw.ops.u8_to_train(dev(frame, torch.uint8)[off:off + n])
        assert bits_equal(got.cpu().numpy(), deploy_ref.u8_to_train(src)).all()
    x = (r.rand(n).astype(np.float32) - 0.5) * 0.999           # in range: the cast is defined
    x[: min(n, 256)] = deploy_ref.u8_to_train(np.arange(256, dtype=np.uint8))[: min(n, 256)]      # every representable level
    got = mgw.ops.train_to_u8(dev(x))
    assert np.array_equal(got.cpu().numpy(), deploy_ref.train_to_u8(x))
    # round trip of every level
    lv = dev(np.arange(256, dtype=np.uint8), torch.uint8)
    back = mgw.ops.train_to_u8(mgw.ops.u8_to_train(lv)).cpu().numpy()
    assert np.array_equal(back, deploy_ref.train_to_u8(deploy_ref.u8_to_train(np.arange(256, dtype=np.uint8))))


def test_fill_zero(mgw):
    for keep in (False, True):
        t = torch.full((3, 97, 64), 7.0, device='cuda')
        mgw.ops.fill_zero(t, keep_in_l2=keep)
        assert float(t.abs().max()) == 0.0
    guard = torch.full((1024 + 8,), 5.0, device='cuda')
    mgw.ops.fill_zero(guard[4:1028], keep_in_l2=True)
    assert float(guard[:4].min()) == 5.0 and float(guard[1028:].min()) == 5.0 and float(guard[4:1028].abs().max()) == 0.0


def test_autograd_reaches_the_fused_backward(mgw, monkeypatch):
    """loss.backward() through transformer_img_loss must take mgw_mesh_warp_img_loss_bwd (d_out formed in registers), and the
    plain transformer must not be handed a materialised all-zero d_img when only `output` is used."""
    from dovs_b200 import functional as F
    calls = {'fused': 0, 'unfused': 0, 'dimg_none': 0}
    fused, unfused = F.ops.mesh_warp_img_loss_bwd, F.ops.mesh_warp_bwd

    def spy_fused(*a, **k):
        calls['fused'] += 1
        return fused(*a, **k)

    def spy_unfused(U, theta, Hs, d_out, d_img=None, **k):
        calls['unfused'] += 1
        calls['dimg_none'] += d_img is None
        return unfused(U, theta, Hs, d_out, d_img, **k)
    monkeypatch.setattr(F.ops, 'mesh_warp_img_loss_bwd', spy_fused)
    monkeypatch.setattr(F.ops, 'mesh_warp_bwd', spy_unfused)
    n, h, w, c = 2, 96, 128, 3
    U, y = dev(synth.smooth_image(n, h, w, c, 590)), dev(synth.smooth_image(n, h, w, c, 591))
    th = dev(synth.random_mesh(n, 4, 4, 0.05, 592)).requires_grad_(True)
    loss, out, black, img = mgw.transformer_img_loss(U, th, y)
    loss.backward()
    assert calls == {'fused': 1, 'unfused': 0, 'dimg_none': 0} and torch.isfinite(th.grad).all()
    th2 = dev(synth.random_mesh(n, 4, 4, 0.05, 592)).requires_grad_(True)
    out2, _, _ = mgw.transformer(U, th2)
    mgw.img_loss(out2, y, black).backward()
    assert calls == {'fused': 1, 'unfused': 1, 'dimg_none': 1}
    assert relmax(th.grad.cpu().numpy(), th2.grad.cpu().numpy()) < 1e-4


# ------------------------------------------------------------------ dtheta as a directional derivative (config #2 size, smooth image)
@pytest.mark.parametrize('impl', ['auto', 'generic'])
def test_dtheta_is_the_directional_derivative(mgw, impl):
    """<dtheta, v> == d/dt L(theta + t v) with L = <out, G> + <img, Gi>, by central differences evaluated with the fp64 oracle
    port (the product computes in fp32: finite differences of ITS outputs would drown in rounding), at config #2's frame size."""
    mgw.set_impl(impl)
    n, h, w, c = 2, 288, 512, 3
    U = synth.smooth_image(n, h, w, c, 600)
    theta = synth.random_mesh(n, 4, 4, 0.05, 601)
    G, Gi = synth.randn((n, h, w, c), 602), synth.randn((n, h, w, 2), 603, 0.1)
    Hs = mgw.ops.solve_h_fwd(dev(theta))
    _, dth = mgw.ops.mesh_warp_bwd(dev(U), dev(theta), Hs, dev(G), dev(Gi))
    v = np.random.RandomState(604).randn(*theta.shape)
    v /= np.linalg.norm(v)

    def L(t):
        o, _, im, _ = ref.transformer(t64(U), t64(theta.astype(np.float64) + t * v))
        return float((o * t64(G)).sum() + (im * t64(Gi)).sum())
    # the objective is only piecewise smooth (border pixels drop in and out of range, see close_but_knife_edges): the
    # step must be small enough that no jump falls inside [-eps, eps]; fp64 leaves room for that
    eps = 1e-6
    fd = (L(eps) - L(-eps)) / (2 * eps)
    got = float((dth.cpu().double().numpy() * v).sum())
    # the fp64 analytic value is the arbiter of what finite differences can resolve
    t = t64(theta, True)
    o, _, im, _ = ref.transformer(t64(U), t)
    ((o * t64(G)).sum() + (im * t64(Gi)).sum()).backward()
    an = float((t.grad.numpy() * v).sum())
    scale = float(np.abs(t.grad.numpy()).max())
    assert abs(fd - an) <= 1e-4 * scale, (fd, an)
    assert abs(got - an) <= 2e-4 * scale, (got, an, fd)


# ------------------------------------------------------------------ multi-GPU: rank-sharded result == single-GPU result
_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'oracle'))
import numpy as np, torch, torch.distributed as dist
import synth, dovs_b200 as mgw
rank, world, local = mgw.parallel.init_from_env('nccl')
dev = torch.device('cuda', local); torch.cuda.set_device(dev)
n, h, w, c = 8, 96, 128, 3
U = torch.tensor(synth.smooth_image(n, h, w, c, 800), device=dev)
y = torch.tensor(synth.smooth_image(n, h, w, c, 801), device=dev)
theta = torch.tensor(synth.random_mesh(n, 4, 4, 0.05, 802), device=dev)
feats = torch.tensor(synth.randn((n, 512), 803), device=dev)
def run(lo, hi):
    Ul, tl = U[lo:hi].contiguous().requires_grad_(True), theta[lo:hi].contiguous().requires_grad_(True)
    loss, out, black, img = mgw.transformer_img_loss(Ul, tl, y[lo:hi].contiguous(), batch_size=n)      # divided by the GLOBAL batch
    loss.backward()
    return loss.detach(), tl.grad, Ul.grad
lo, hi = mgw.parallel.shard_bounds(n, rank, world)
loss_l, dth_l, dU_l = run(lo, hi)
red = mgw.parallel.MeshHeadGradReducer(512, 50, dev)
red.launch(0, features=feats[lo:hi].contiguous(), dtheta=dth_l)
head = red.wait().clone()
dist.all_reduce(loss_l)
gath = [torch.empty_like(dth_l) for _ in range(world)]
dist.all_gather(gath, dth_l)
torch.cuda.synchronize()
if rank == 0:
    loss_1, dth_1, dU_1 = run(0, n)
    red1 = mgw.parallel.MeshHeadGradReducer(512, 50, dev)
    head_1 = red1.head_grad(feats, dth_1).clone()
    dth_s = torch.cat(gath)
    assert torch.equal(dth_s, dth_1), 'sharded dtheta differs: %%g' %% float((dth_s - dth_1).abs().max())
    assert torch.equal(dU_l, dU_1[lo:hi])
    assert abs(float(loss_l) - float(loss_1)) <= 1e-6 * abs(float(loss_1))
    e = float((head - head_1).abs().max() / head_1.abs().max())
    assert e < 1e-5, e
    print('MULTI_GPU_OK', world, e)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_gpu_sharded_step_equals_single_gpu(mgw, tmp_path):
    """2 ranks (one per GPU, NCCL): the batch-sharded fused warp+img_loss backward gives bit-identical per-sample dtheta / dU,
    and the all-reduced mesh-head gradient equals the single-GPU one (train_bundle_nobm.py's step under data parallelism)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER % {'root': ROOT})
    env = dict(os.environ)
    env.pop('CUDA_VISIBLE_DEVICES', None)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29731', str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and 'MULTI_GPU_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_double_buffered_dU_pipeline(mgw):
    """the usage bench.py times: dU double-buffered, step i accumulates into buffer i % 2 (mgw_mesh_warp_bwd_acc) while the zero-fill
    of the other buffer -- the one step i+1 accumulates into -- runs next to it on a side stream.  Every step's dU and dtheta equal
    the plain call's (mgw_mesh_warp_bwd zero-fills inside), over several steps with different inputs, eager and as CUDA graphs."""
    mgw.set_impl('auto')
    n, h, w, c = 4, 96, 128, 3
    sets = [dict(U=dev(synth.noise_image(n, h, w, c, 600 + k)), th=dev(synth.random_mesh(n, 4, 4, 0.05, 610 + k)),
                 g=dev(synth.randn((n, h, w, c), 620 + k)), gi=dev(synth.randn((n, h, w, 2), 630 + k, 0.1))) for k in range(3)]
    want = []
    for s in sets:
        _, _, _, Hs = mgw.ops.mesh_warp_fwd(s['U'], s['th'])
        dU, dth = mgw.ops.mesh_warp_bwd(s['U'], s['th'], Hs, s['g'], s['gi'])
        want.append((dU.clone(), dth.clone()))
    bufs = [torch.zeros_like(sets[0]['U']) for _ in range(2)]
    dths = [torch.empty_like(sets[0]['th']) for _ in range(3)]
    side = torch.cuda.Stream()

    def step(i):
        s = sets[i % 3]
        cur = torch.cuda.current_stream()
        _, _, _, Hs = mgw.ops.mesh_warp_fwd(s['U'], s['th'])
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            mgw.ops.fill_zero(bufs[(i + 1) % 2])
        mgw.ops.mesh_warp_bwd(s['U'], s['th'], Hs, s['g'], s['gi'], accumulate_into=bufs[i % 2], dtheta_out=dths[i % 3])
        cur.wait_stream(side)

    def check_step(i):
        torch.cuda.synchronize()
        dU, dth = want[i % 3]
        # same kernels, same data: the reductions of a tile arrive in any order, so dU agrees to fp32 summation noise
        assert relmax(bufs[i % 2].cpu().numpy(), dU.cpu().numpy()) <= 2e-6, i
        assert torch.equal(dths[i % 3], dth), i

    for i in range(6):
        step(i)
        check_step(i)
    s_cap = torch.cuda.Stream()
    s_cap.wait_stream(torch.cuda.current_stream())
    graphs = []
    with torch.cuda.stream(s_cap):
        for i in range(6):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s_cap):
                step(i)
            graphs.append(g)
    torch.cuda.current_stream().wait_stream(s_cap)
    for rep in range(2):
        for i in range(6):
            graphs[i].replay()
            check_step(i)


@pytest.mark.parametrize('sigma', [0.05, 0.3])
def test_backward_next_to_its_own_zero_fill(mgw, sigma):
    """mgw_mesh_warp_bwd zero-fills dU itself and launches the tile backward PROGRAMMATICALLY behind that fill: the kernel runs next to
    the fill up to its first access to dU (griddepcontrol.wait before the drain; before the loop on the general path, which
    sigma = 0.3 exercises).  At config #2's size the fill takes ~10 us, so a missing wait would let reductions land in a buffer that
    is zeroed afterwards or still holds garbage: dU is pre-filled with NaN before every call and must come out equal to the
    accumulate-into-zeros form (no fill inside the call, plain launch), eager back to back and as a replayed CUDA graph.
    (Checked once by mutation: a build WITHOUT the two waits passes this test as well -- the fill sweeps the buffer in address order
    and stays ahead of the tiles, which run in the same order -- so the test guards the call sequence and the NaN-prefilled buffer,
    while the guarantee itself is the griddepcontrol.wait in mgw_warp_tma.cu.)"""
    mgw.set_impl('auto')
    n, h, w, c = 32, 288, 512, 3
    U = dev(synth.noise_image(n, h, w, c, 700)); th = dev(synth.random_mesh(n, 4, 4, sigma, 701))
    g = dev(synth.randn((n, h, w, c), 702)); gi = dev(synth.randn((n, h, w, 2), 703, 0.1))
    _, _, _, Hs = mgw.ops.mesh_warp_fwd(U, th)
    ref = torch.zeros_like(U)
    _, dth_ref = mgw.ops.mesh_warp_bwd(U, th, Hs, g, gi, accumulate_into=ref)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    buf = torch.empty_like(U)

    def run():
        buf.fill_(float('nan'))
        return mgw.ops.mesh_warp_bwd(U, th, Hs, g, gi, dU_out=buf)

    def check(dth):
        torch.cuda.synchronize()
        assert torch.isfinite(buf).all()
        assert float((buf - ref).abs().max()) <= 2e-6 * scale        # same kernels, same data: fp32 order of the tiles' reductions only
        assert torch.equal(dth, dth_ref)

    for _ in range(4):
        _, dth = run()
        check(dth)
    s_cap = torch.cuda.Stream()
    s_cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_cap):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s_cap):
            _, dth_g = run()
    torch.cuda.current_stream().wait_stream(s_cap)
    for _ in range(4):
        gr.replay()
        check(dth_g)


@pytest.mark.parametrize('n,gh,gw', [(3, 12, 12), (2, 16, 20), (5, 1, 1), (33, 3, 7)])
def test_solve_kernels_on_large_and_odd_grids(mgw, n, gh, gw):
    """K1 / K4 beyond the 4 x 4 mesh: more cells per sample than K4's block has threads (its per-cell loop runs more than once),
    cell counts that are not multiples of K1's 16 cells per block, a single cell.  K1 bit-identical to the C oracle; K4 (from given
    dHs, and from per-tile partials with a count that is not a multiple of four through the warp backward) against fp64 autograd
    of the oracle's solve (spatial_transformer3.py:144-198)."""
    theta = synth.random_mesh(n, gh, gw, 0.15 / max(gh, gw), 7000 + gh)
    Hs = mgw.ops.solve_h_fwd(dev(theta))
    assert bits_equal(Hs.cpu().numpy(), c_oracle.solve_h(theta)).all()
    dHs = synth.randn((n, gh, gw, 9), 7100 + gw)
    th64 = t64(theta, grad=True)
    H64 = ref.solve_h(th64)
    up = t64(dHs).clone()
    up[..., 8] = 0
    (H64 * up).sum().backward()
    # K4 reads h6, h7 from Hs: with the fp64 solution (rounded to fp32) it reproduces the fp64 adjoint; with K1's own fp32 Hs the
    # small cells of a fine mesh (an ill-conditioned 8x8 system) carry K1's error into dtheta, as the reference's fp32 graph does
    got = mgw.ops.solve_h_bwd(dev(theta), dev(H64.detach().numpy()), dev(dHs)).cpu().numpy()
    assert relmax(got, th64.grad.numpy()) < 2e-5, relmax(got, th64.grad.numpy())
    got32 = mgw.ops.solve_h_bwd(dev(theta), Hs, dev(dHs)).cpu().numpy()
    assert relmax(got32, th64.grad.numpy()) < (2e-5 if gh * gw <= 25 else 1e-2)
    # through the warp backward: tile partials (TMA path when the cells are large enough, generic otherwise) -> K4
    h, w = gh * 24, gw * 32
    if n * h * w <= 4 * 288 * 512:
        U = dev(synth.smooth_image(n, h, w, 1, 7200))
        g = dev(synth.randn((n, h, w, 1), 7300))
        _, dHs2 = mgw.ops.warp_bwd(U, Hs, g, want_dU=False)
        _, dth = mgw.ops.mesh_warp_bwd(U, dev(theta), Hs, g, want_dU=False)
        want = mgw.ops.solve_h_bwd(dev(theta), Hs, dHs2)
        assert relmax(dth.cpu().numpy(), want.cpu().numpy()) < 1e-5
