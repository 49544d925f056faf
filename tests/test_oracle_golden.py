"""CPU: pin both oracles (oracle/mgw_oracle.c, oracle/mesh_warp_ref.py) to the golden vectors, which are
outputs of the unmodified reference sources (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

import c_oracle
import mesh_warp_ref as ref
from conftest import (MESH_CASES, SMALL_MESH_CASES, bits_equal, golden_black, golden_inputs, load_golden, relmax)


@pytest.mark.parametrize('name', MESH_CASES)
def test_c_oracle_pixel_stage_bit_exact(name):
    """given the reference's Hs: x_map/y_map, black_pix, output_img reproduce the reference BIT FOR BIT."""
    g = load_golden(name)
    U, _, _ = golden_inputs(name, g)
    out, black, img, cell = c_oracle.warp(U, g['ref_Hs'])
    assert bits_equal(img, g['ref_img']).all()
    assert (black == golden_black(g)).all()
    assert bits_equal(out, g['ref_out']).all()
    gh, gw = g['grid']
    assert cell.min() == 0 and cell.max() == gh * gw - 1


@pytest.mark.parametrize('name', MESH_CASES)
def test_c_oracle_solve_h(name):
    """the 8x8 solve is LAPACK-shaped, not LAPACK: |dH| <= 1e-5 abs and no worse vs fp64 than 2x the reference."""
    g = load_golden(name)
    Hc = c_oracle.solve_h(g['theta'])
    assert np.abs(Hc - g['ref_Hs']).max() < 1e-5
    assert (Hc[..., 8] == 1).all()
    if 'f64_Hs' in g:
        H64 = c_oracle.solve_h(g['theta'], f64=True)
        assert np.abs(H64 - g['f64_Hs']).max() < 1e-6
        assert np.abs(Hc - g['f64_Hs']).max() <= 2 * np.abs(g['ref_Hs'] - g['f64_Hs']).max() + 1e-6


@pytest.mark.parametrize('name', SMALL_MESH_CASES)
def test_torch_port_forward_and_grads(name):
    g = load_golden(name)
    U, d_out, d_img = golden_inputs(name, g)
    Ut = torch.tensor(U, requires_grad=True)
    th = torch.tensor(g['theta'], requires_grad=True)
    out, black, img, Hs = ref.transformer(Ut, th)
    Hs.retain_grad()
    ((out * torch.tensor(d_out)).sum() + (img * torch.tensor(d_img)).sum()).backward()
    # same LAPACK behind torch.linalg.inv as the shim used -> Hs identical, everything downstream ~bit-exact
    assert np.abs(Hs.detach().numpy() - g['ref_Hs']).max() < 1e-6
    fin = np.isfinite(g['ref_out']).all(-1) & (np.abs(g['ref_img']).max(-1) < 1e3)
    assert (black.numpy() == golden_black(g))[fin].all()
    assert np.abs(out.detach().numpy() - g['ref_out'])[fin].max() < 1e-5
    if name != 'mesh_fold_clamp_shift':       # a folded cell has z -> 0 pixels: gradients there are garbage in the reference too
        assert relmax(Ut.grad.numpy(), g['ref_dU']) < 1e-4
        assert relmax(th.grad.numpy(), g['ref_dtheta']) < 1e-4
        assert relmax(Hs.grad.numpy()[..., :8], g['ref_dHs'][..., :8]) < 1e-4


@pytest.mark.parametrize('name', SMALL_MESH_CASES[:5])
def test_torch_port_fp64_arbiter(name):
    g = load_golden(name)
    U, d_out, d_img = golden_inputs(name, g)
    Ut = torch.tensor(U, dtype=torch.float64, requires_grad=True)
    th = torch.tensor(g['theta'], dtype=torch.float64, requires_grad=True)
    out, black, img, Hs = ref.transformer(Ut, th)
    ((out * torch.tensor(d_out, dtype=torch.float64)).sum() + (img * torch.tensor(d_img, dtype=torch.float64)).sum()).backward()
    assert np.abs(Hs.detach().numpy() - g['f64_Hs']).max() < 1e-9
    assert relmax(th.grad.numpy(), g['f64_dtheta']) < 1e-8
    assert relmax(Ut.grad.numpy(), g['f64_dU']) < 1e-8


@pytest.mark.parametrize('name', ['homog_s05', 'homog_c1_s10'])
def test_homography_oracles(name):
    g = load_golden(name)
    n, h, w, c = g['U'].shape
    out, black, _ = c_oracle.homography_warp(g['U'], g['theta'], (h, w))
    assert bits_equal(out, g['ref_out']).all()
    assert (black == g['ref_black']).all()
    Ut = torch.tensor(g['U'], requires_grad=True)
    th = torch.tensor(g['theta'], requires_grad=True)
    o, b = ref.transformer_homography(Ut, th, (h, w))
    (o * torch.tensor(g['d_out'])).sum().backward()
    assert np.abs(o.detach().numpy() - g['ref_out']).max() < 1e-6
    assert relmax(Ut.grad.numpy(), g['ref_dU']) < 1e-5
    assert relmax(th.grad.numpy(), g['ref_dtheta']) < 1e-4


@pytest.mark.parametrize('name', ['interp_same', 'interp_resize_c1'])
def test_interp_oracles(name):
    g = load_golden(name)
    oh, ow = g['x'].shape[1:3]
    out = c_oracle.interp(g['im'], g['x'], g['y'], (oh, ow))
    assert bits_equal(out, g['ref_out']).all()
    it = torch.tensor(g['im'], requires_grad=True)
    xt = torch.tensor(g['x'], requires_grad=True)
    yt = torch.tensor(g['y'], requires_grad=True)
    o = ref.interpolate(it, xt, yt, (oh, ow))
    (o * torch.tensor(g['d_out'])).sum().backward()
    assert bits_equal(o.detach().numpy(), g['ref_out']).all()
    assert relmax(it.grad.numpy(), g['ref_dim']) < 1e-5
    assert relmax(xt.grad.numpy(), g['ref_dx']) < 1e-5
    assert relmax(yt.grad.numpy(), g['ref_dy']) < 1e-5


def test_losses_and_vertices_oracles():
    g = load_golden('losses')
    gh, gw = (int(v) for v in g['grid'])
    n, h, w, _ = g['x1'].shape
    p1, p2 = c_oracle.vertices(g['head'], gh, gw)
    assert bits_equal(p1, g['ref_pts1']).all() and bits_equal(p2, g['ref_pts2']).all()
    hd = torch.tensor(g['head'], requires_grad=True)
    hd2 = torch.tensor(g['head2'], requires_grad=True)
    pts1, pts2 = ref.get_4_pts(hd, gh, gw)
    _, pts2b = ref.get_4_pts(hd2, gh, gw)
    assert bits_equal(pts1.detach().numpy(), g['ref_pts1']).all()
    out1, black1, img1, _ = ref.transformer(torch.tensor(g['x1']), pts2)
    out2, black2, img2, _ = ref.transformer(torch.tensor(g['x2']), pts2b)
    fl, warpped = ref.feature_loss(torch.tensor(g['matches']), torch.tensor(g['mask']), img1)
    il = ref.img_loss(out1, torch.tensor(g['y1']), black1)
    tl = ref.temp_loss(out1, black1, out2, black2, torch.tensor(g['flow']))
    for nm, l in (('feature', fl), ('img', il), ('temp', tl)):
        assert abs(float(l) - float(g['ref_%s_loss' % nm])) <= 1e-5 * max(1.0, abs(float(g['ref_%s_loss' % nm]))), nm
        gr = torch.autograd.grad(l, [hd, hd2], retain_graph=True, allow_unused=True)
        for k, t in (('dhead', gr[0]), ('dhead2', gr[1])):
            want = g['ref_%s_%s' % (k, nm)]
            got = np.zeros_like(want) if t is None else t.numpy()
            if np.abs(want).max() > 0:
                assert relmax(got, want) < 2e-4, (nm, k)
            else:
                assert np.abs(got).max() == 0


@pytest.mark.reference
def test_golden_is_reference_output():
    """regenerate one fixture from /root/reference and compare with the committed file."""
    import make_golden
    import synth
    name, fx = make_golden.mesh_case('mesh_smooth_s03', synth.smooth_image(2, 48, 64, 3, 10),
                                     synth.random_mesh(2, 4, 4, 0.03, 11), 4, 4, 10, f64=None)
    g = load_golden(name)
    for k in ('ref_Hs', 'ref_out', 'ref_img', 'ref_dtheta', 'ref_dU'):
        assert bits_equal(fx[k], g[k]).all(), k


# ------------------------------------------------------------------ deploy side: warpRevBundle2 (deploy_bundle.py:136-146)
DEPLOY_CASES = ['net', 'ragged', 'identity']


@pytest.mark.parametrize('tag', DEPLOY_CASES)
def test_deploy_oracle_matches_the_reference_output(tag):
    """oracle/deploy_ref.py (numpy) == the reference's warpRevBundle2 run on OpenCV (plain C++ path), byte for byte."""
    import deploy_ref
    g = load_golden('deploy_remap')
    dst = deploy_ref.warp_rev_bundle2(g[tag + '_img'], g[tag + '_x_map'], g[tag + '_y_map'])
    assert np.array_equal(dst, g[tag + '_ref_dst'])
    # what OpenCV's SIMD-dispatched build does to the same call (recorded when the fixture was made): a handful of bytes
    assert float(g[tag + '_opt_diff_frac']) < 1e-3


@pytest.mark.reference
def test_deploy_golden_is_reference_output():
    """regenerates one deploy fixture from /root/reference (needs cv2) and compares with the committed one."""
    cv2 = pytest.importorskip('cv2')
    import ref_loader
    g = load_golden('deploy_remap')
    h, w = g['ragged_x_map'].shape
    fn = ref_loader.deploy_warp_rev_bundle2(h, w)
    cv2.setUseOptimized(False)
    try:
        dst = fn(g['ragged_img'], g['ragged_x_map'].copy(), g['ragged_y_map'].copy())
    finally:
        cv2.setUseOptimized(True)
    assert np.array_equal(dst, g['ragged_ref_dst'])


# ------------------------------------------------------------------ deploy side: streaming state (deploy_bundle.py:204-328)
def test_stream_oracle_matches_the_reference_statements():
    """oracle/deploy_ref.StreamStateRef == the reference's own list handling (its statements exec'd, oracle/make_golden.py)."""
    import deploy_ref
    g = load_golden('deploy_stream')
    st = deploy_ref.StreamStateRef(g['first'])
    for k in range(g['cur_frames'].shape[0]):
        in_x = st.assemble(g['cur_frames'][k])
        assert np.array_equal(in_x, g['in_x'][k])
        tmp = in_x.copy()
        for _ in range(int(g['refine'])):
            img, black = deploy_ref.stream_fake_net(tmp, k)
            tmp[0, :, :, -1] = st.frame_of(img, black)
        assert np.array_equal(tmp, g['tmp_in_x'][k])
        st.push(img, black)
    assert np.array_equal(np.concatenate(st.frames, 0), g['final_frames'])
    assert np.array_equal(np.concatenate(st.masks, 0), g['final_masks'])


@pytest.mark.reference
def test_stream_golden_is_reference_output():
    """the reference's statements are still where ref_loader picks them (line ranges of deploy_bundle.py)"""
    import ref_loader
    blocks = ref_loader.deploy_stream_blocks()
    assert set(blocks) == {'assemble', 'refine', 'update'}


def test_crop_oracle_matches_the_reference_loops():
    """deploy_ref.crop_rect (run-length restatement) == `ans` of deploy_bundle.py:344-365 executed verbatim (fixture)"""
    import deploy_ref
    g = load_golden('deploy_crop')
    names = sorted(k[:-4] for k in g if k.endswith('_ans'))
    assert len(names) >= 6
    for name in names:
        ans = deploy_ref.crop_rect(g[name + '_all_black'])
        assert ans == g[name + '_ans'].tolist(), name
        if ans:
            assert (ans[2] - ans[0] + 1) * (ans[3] - ans[1] + 1) == int(g[name + '_max_s'])
            assert g[name + '_all_black'][ans[0]:ans[2] + 1, ans[1]:ans[3] + 1].sum() == 0
    s = load_golden('deploy_stream')
    ab = np.zeros(s['first'].shape, np.int64)
    ref = deploy_ref.StreamStateRef(s['first'])
    for k in range(s['cur_frames'].shape[0]):
        in_x = ref.assemble(s['cur_frames'][k])
        for _ in range(int(s['refine'])):
            img, black = deploy_ref.stream_fake_net(in_x, k)
            ab = deploy_ref.black_accumulate(ab, black)
            in_x[..., -1] = ref.frame_of(img, black)
        ref.push(img, black)
    assert np.array_equal(ab, s['all_black'])


@pytest.mark.reference
def test_crop_golden_is_reference_output():
    """regenerate one crop case from the reference's own statements"""
    import math
    import ref_loader
    g = load_golden('deploy_crop')
    ab = g['tie_all_black']
    ns = dict(np=np, math=math, height=ab.shape[0], width=ab.shape[1], all_black=ab)
    ref_loader.quiet(exec, ref_loader.deploy_crop_block(), ns)
    assert ns['ans'] == g['tie_ans'].tolist() and ns['max_s'] == int(g['tie_max_s'])
    # and the restatement against the reference's loops on random masks (sparse and dense black, ragged sizes)
    import deploy_ref
    code = ref_loader.deploy_crop_block()
    r = np.random.RandomState(17)
    for k in range(40):
        h, w = int(r.randint(12, 48)), int(r.randint(12, 64))
        ab = (r.random_sample((h, w)) < r.choice([0.0, 0.01, 0.05, 0.3])).astype(np.int64) * r.randint(1, 5, (h, w))
        ns = dict(np=np, math=math, height=h, width=w, all_black=ab)
        ref_loader.quiet(exec, code, ns)
        assert deploy_ref.crop_rect(ab) == ns['ans'], (k, h, w)


def test_vertex_loss_oracle_matches_the_reference_functions():
    """oracle/vertex_loss_ref.py (closed forms, values + gradients) == the reference's get_black_pos / get_distortion_loss /
    get_consistency_loss / id loss run on the shim (fixture), fp32 to rounding and fp64 to 1e-12"""
    import vertex_loss_ref as V
    g = load_golden('vertex_losses')

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) if np.abs(b).max() > 0 else np.abs(a).max()

    assert len(g['grids']) >= 5
    for gi, (gh, gw) in enumerate(g['grids']):
        for tag, dt, tol in (('ref', np.float32, 2e-6), ('f64', np.float64, 1e-12)):
            k = 'g%d_%s_' % (gi, tag)
            head, raw = g['g%d_head' % gi].astype(dt), g['g%d_raw' % gi].astype(dt)
            p1, p2 = g[k + 'pts1'].astype(dt), g[k + 'pts2'].astype(dt)
            v, gr = V.id_loss(head)
            assert rel(v, g[k + 'id']) <= tol and rel(gr, g[k + 'did_dhead']) <= tol
            v, gr, err = V.black_pos_loss(raw)
            assert rel(v, g[k + 'black']) <= tol and rel(gr, g[k + 'dblack_draw']) <= tol
            assert rel(err.reshape(err.shape[0], -1), g[k + 'black_err']) <= tol
            assert rel((err * err).reshape(err.shape[0], -1), g[k + 'black_pos']) <= tol
            assert V.black_pos_loss(p1)[0] == 0 and g[k + 'black_clamped'] == 0
            v, gr = V.distortion_loss(p1, int(gh), int(gw))
            assert rel(v, g[k + 'dist']) <= tol and rel(gr, g[k + 'ddist_dpts1']) <= tol, (gh, gw, tag)
            v, gr = V.consistency_loss(p2, int(gh), int(gw))
            assert rel(v, g[k + 'cons']) <= tol and rel(gr, g[k + 'dcons_dpts2']) <= tol, (gh, gw, tag)


@pytest.mark.reference
def test_vertex_loss_golden_is_reference_output():
    """regenerate one value from the reference's own function"""
    import ref_loader
    import torch
    g = load_golden('vertex_losses')
    gh, gw = (int(v) for v in g['grids'][1])
    ns = ref_loader.s_net_regularisers(ref_loader.s_net_namespace(8, 8, gh, gw, 3, 1))
    pts1, pts2 = ns['get_4_pts'](torch.as_tensor(g['g1_head']), 3)
    assert np.allclose(ref_loader.quiet(ns['get_distortion_loss'], pts1).numpy(), g['g1_ref_dist'], rtol=1e-6)
    assert np.allclose(ref_loader.quiet(ns['get_consistency_loss'], pts2).numpy(), g['g1_ref_cons'], rtol=1e-6)


def test_warp_rev_bundle_oracle_matches_the_reference():
    """deploy_ref.warp_rev_bundle (cv2.warpPerspective restated: double-precision block walk, 1/32-px fixed point) ==
    warpRevBundle of deploy_bundle.py:148-173 run from the reference source on OpenCV (fixture), byte for byte"""
    import deploy_ref
    g = load_golden('deploy_warp_rev_bundle')
    names = sorted(k[:-4] for k in g if k.endswith('_dst'))
    assert len(names) >= 5
    for n in names:
        gh, gw = (int(v) for v in g[n + '_grid'])
        h, w = g[n + '_img'].shape[:2]
        assert np.array_equal(deploy_ref.cvt_theta_mat_bundle(g[n + '_Hs'], h, w, gh, gw), g[n + '_Hs_cvt']), n
        assert np.array_equal(deploy_ref.warp_rev_bundle(g[n + '_img'], g[n + '_Hs'], gh, gw), g[n + '_dst']), n
        assert int(g[n + '_optimized_differs']) == 0          # OpenCV's two code paths agree on this one
    assert np.array_equal(deploy_ref.warp_rev_bundle(g['strong_img'], g['warprev_theta'].reshape(1, 1, 9), 1, 1), g['warprev_out'])


@pytest.mark.reference
def test_warp_rev_bundle_golden_is_reference_output():
    cv2 = pytest.importorskip('cv2')
    import ref_loader
    g = load_golden('deploy_warp_rev_bundle')
    gh, gw = (int(v) for v in g['ragged_grid'])
    h, w = g['ragged_img'].shape[:2]
    warp, _ = ref_loader.deploy_warp_rev_bundle(h, w, gh, gw)
    cv2.setUseOptimized(False)
    try:
        assert np.array_equal(warp(g['ragged_img'], g['ragged_Hs']), g['ragged_dst'])
    finally:
        cv2.setUseOptimized(True)


def test_cvt_img2train_oracle_matches_the_reference():
    """deploy_ref.cvt_img2train (cv2 BGR2GRAY + Pillow BILINEAR resize restated) == config.py:6-21 run from the reference
    source on OpenCV and Pillow (fixture), exactly (float64 values)"""
    import deploy_ref
    g = load_golden('deploy_cvt_img2train')
    names = sorted(k[:-4] for k in g if k.endswith('_out'))
    assert len(names) >= 6
    for n in names:
        h, w, cr = g[n + '_cfg']
        got = deploy_ref.cvt_img2train(g[n + '_img'], int(h), int(w), 1 if cr == 1 else float(cr))
        assert got.dtype == np.float64 and np.array_equal(got, g[n + '_out']), n


@pytest.mark.reference
def test_cvt_img2train_golden_is_reference_output():
    pytest.importorskip('cv2'); pytest.importorskip('PIL')
    import ref_loader
    g = load_golden('deploy_cvt_img2train')
    f = ref_loader.config_cvt_img2train(72, 128)
    assert np.array_equal(f(g['crop_img'], 0.9), g['crop_out'])


def test_cv2_resize_oracle_matches_opencv():
    """deploy_ref.resize_linear_u8 == cv2.resize(uint8 frame, dsize) (fixture made by OpenCV itself), byte for byte"""
    import deploy_ref
    g = load_golden('deploy_cv2_resize')
    names = sorted(k[:-4] for k in g if k.endswith('_dst'))
    assert len(names) >= 6
    for n in names:
        want = g[n + '_dst']
        assert np.array_equal(deploy_ref.resize_linear_u8(g[n + '_img'], want.shape[1], want.shape[0]), want), n
        assert int(g[n + '_optimized_differs']) == 0


@pytest.mark.reference
def test_deploy_oracles_against_the_reference_on_random_inputs():
    """beyond the committed fixtures: the OpenCV / Pillow restatements against the reference's own functions (and the libraries
    they call) on random frames, sizes and homographies -- only where /root/reference, cv2 and Pillow exist"""
    cv2 = pytest.importorskip('cv2'); pytest.importorskip('PIL')
    import deploy_ref
    import ref_loader
    r = np.random.RandomState(23)
    cv2.setUseOptimized(False)
    try:
        for k in range(6):
            H, W = int(r.randint(40, 200)), int(r.randint(40, 260))
            h, w = int(r.randint(16, 120)), int(r.randint(16, 160))
            img = r.randint(0, 256, (H, W, 3)).astype(np.uint8)
            cr = float(r.choice([1.0, 0.9, 0.75]))
            f = ref_loader.config_cvt_img2train(h, w)
            want = f(img, cr) if cr != 1 else f(img)
            assert np.array_equal(deploy_ref.cvt_img2train(img, h, w, 1 if cr == 1 else cr), want), ('cvt', H, W, h, w, cr)
            assert np.array_equal(deploy_ref.resize_linear_u8(img, w, h), cv2.resize(img, (w, h))), ('resize', H, W, h, w)
            gh, gw = int(r.randint(1, 5)), int(r.randint(1, 5))
            Hs = (np.tile(np.eye(3, dtype=np.float32).reshape(1, 1, 9), (gh, gw, 1)) +
                  0.1 * r.standard_normal((gh, gw, 9)).astype(np.float32) * np.array([1, 1, 1, 1, 1, 1, 0.5, 0.5, 0], np.float32)).astype(np.float32)
            warp, _ = ref_loader.deploy_warp_rev_bundle(H, W, gh, gw)
            assert np.array_equal(deploy_ref.warp_rev_bundle(img, Hs, gh, gw), warp(img, Hs)), ('warpRevBundle', H, W, gh, gw)
            x_map = (np.linspace(-1, 1, W, dtype=np.float32)[None, :] + 0.05 * r.standard_normal((H, W)).astype(np.float32)).astype(np.float32)
            y_map = (np.linspace(-1, 1, H, dtype=np.float32)[:, None] + 0.05 * r.standard_normal((H, W)).astype(np.float32)).astype(np.float32)
            assert np.array_equal(deploy_ref.warp_rev_bundle2(img, x_map, y_map), ref_loader.deploy_warp_rev_bundle2(H, W)(img, x_map, y_map)), ('wrb2', H, W)
    finally:
        cv2.setUseOptimized(True)
