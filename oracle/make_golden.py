"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by EXECUTING THE REFERENCE SOURCES.

Run inside the build container only (needs /root/reference):

    python oracle/make_golden.py            # rewrites every fixture
    python oracle/make_golden.py mesh_fold  # one case

The reference has no tests or golden vectors of its own, so every
known-answer vector here is an output of /root/reference/spatial_transformer3.py,
spatial_transformer.py, s_net_bundle_nobm.py (get_4_pts, warp_pts, feature_loss,
img_loss) and train_bundle_nobm.py (temp_loss), run unmodified on the torch-CPU
`tensorflow` shim (TensorFlow is not installable here).  Each fixture carries

  * the inputs,
  * the fp32 outputs and autograd gradients of the reference ("ref_*"),
  * for small cases the same source run in fp64 ("f64_*") as the arbiter.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader as rl      # noqa: E402
import synth                 # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy()


def _tt(a, dt, grad=False):
    t = torch.tensor(np.asarray(a), dtype=dt)
    if grad:
        t.requires_grad_(True)
    return t


def run_mesh(U, theta, d_out, d_img, gh, gw, dt=torch.float32, want_grads=True):
    """reference spatial_transformer3.transformer + autograd of <out,d_out> + <img,d_img>."""
    tf = rl.tf()
    tf.DEFAULT_FLOAT[0] = dt
    try:
        n, h, w, c = U.shape
        m = rl.st3(h, w, gh, gw)
        Ut, th = _tt(U, dt, want_grads), _tt(theta, dt, want_grads)
        tf.NAMED.clear()
        out, black, img = rl.quiet(m.transformer, Ut, th)
        Hs = tf.NAMED['Hs']
        res = dict(Hs=_np(Hs), out=_np(out), black=_np(black), img=_np(img))
        assert tf.NAMED['x_map'].shape == (n, h, w, 1) and tf.NAMED['output_img'].shape == (n, h, w, c)
        if want_grads:
            Hs.retain_grad()
            loss = (out * _tt(d_out, dt)).sum() + (img * _tt(d_img, dt)).sum()
            loss.backward()
            res.update(dU=_np(Ut.grad), dtheta=_np(th.grad), dHs=_np(Hs.grad))
        return res
    finally:
        tf.DEFAULT_FLOAT[0] = torch.float32


def mesh_case(name, U, theta, gh, gw, seed, f64=True, store_inputs=True, store_dU=True):
    n, h, w, c = U.shape
    d_out = synth.randn((n, h, w, c), seed + 100)
    d_img = synth.randn((n, h, w, 2), seed + 200, 0.1)
    ref = run_mesh(U, theta, d_out, d_img, gh, gw)
    fx = dict(theta=theta, grid=np.array([gh, gw]), shape=np.array([n, h, w, c]), seed=np.array(seed))
    if store_inputs:
        fx.update(U=U, d_out=d_out, d_img=d_img)
    for k, v in ref.items():
        if k == 'black':
            assert set(np.unique(v)).issubset({0.0, 1.0})
            fx['ref_black_bits'] = np.packbits(v.astype(np.uint8).reshape(-1))
        elif k == 'dU' and not store_dU:
            continue
        else:
            fx['ref_' + k] = v
    if f64:
        r64 = run_mesh(U, theta, d_out, d_img, gh, gw, dt=torch.float64)
        for k in ('Hs', 'dtheta', 'dHs', 'dU', 'out', 'img'):
            if k == 'dU' and not store_dU:
                continue
            fx['f64_' + k] = r64[k]
    return name, fx


def gen_mesh_cases():
    H, W = 48, 64
    yield mesh_case('mesh_smooth_s03', synth.smooth_image(2, H, W, 3, 10), synth.random_mesh(2, 4, 4, 0.03, 11), 4, 4, 10)
    yield mesh_case('mesh_noise_s08', synth.noise_image(2, H, W, 3, 20), synth.random_mesh(2, 4, 4, 0.08, 21), 4, 4, 20)
    yield mesh_case('mesh_identity', synth.smooth_image(1, H, W, 3, 30), synth.identity_mesh(1, 4, 4), 4, 4, 30)
    yield mesh_case('mesh_ragged_c1', synth.noise_image(2, 50, 70, 1, 40), synth.random_mesh(2, 4, 4, 0.05, 41), 4, 4, 40)
    yield mesh_case('mesh_grid23_c4', synth.smooth_image(1, 40, 60, 4, 50), synth.random_mesh(1, 2, 3, 0.05, 51), 2, 3, 50)
    # adversarial: folded cell (two vertices swapped), vertices pinned at +-1.25, pure translation by 0.5
    th = synth.identity_mesh(3, 4, 4)
    th[0, 2, 2], th[0, 2, 3] = th[0, 2, 3].copy(), th[0, 2, 2].copy()
    th[1] = np.clip(th[1] * 1.4, -1.25, 1.25)
    th[2, ..., 0] += 0.5
    yield mesh_case('mesh_fold_clamp_shift', synth.smooth_image(3, H, W, 3, 60), th, 4, 4, 60)
    # full-size (config #1 / #2 shape, one sample): inputs are regenerated from synth seeds, dU not stored
    yield mesh_case('mesh_full_noise_s05', synth.noise_image(1, 288, 512, 3, 70), synth.random_mesh(1, 4, 4, 0.05, 71),
                    4, 4, 70, f64=False, store_inputs=False, store_dU=False)
    yield mesh_case('mesh_full_identity', synth.smooth_image(1, 288, 512, 3, 80), synth.identity_mesh(1, 4, 4),
                    4, 4, 80, f64=False, store_inputs=False, store_dU=False)


def gen_homography_cases():
    """spatial_transformer.transformer(U, theta[N,9], out_size) (spatial_transformer.py:18-197)."""
    m = rl.st1()
    for name, n, h, w, c, sigma, seed in (('homog_s05', 2, 48, 64, 3, 0.05, 110), ('homog_c1_s10', 3, 30, 44, 1, 0.10, 120)):
        U = synth.smooth_image(n, h, w, c, seed)
        th = synth.random_homography(n, sigma, seed + 1)
        d_out = synth.randn((n, h, w, c), seed + 2)
        fx = dict(U=U, theta=th, d_out=d_out)
        for tag, dt in (('ref', torch.float32), ('f64', torch.float64)):
            tf = rl.tf()
            tf.DEFAULT_FLOAT[0] = dt
            try:
                Ut, tt = _tt(U, dt, True), _tt(th, dt, True)
                out, black = rl.quiet(m.transformer, Ut, tt, (h, w))
                (out * _tt(d_out, dt)).sum().backward()
                fx.update({tag + '_out': _np(out), tag + '_black': _np(black), tag + '_dU': _np(Ut.grad),
                           tag + '_dtheta': _np(tt.grad)})
            finally:
                tf.DEFAULT_FLOAT[0] = torch.float32
        yield name, fx


def gen_interp_cases():
    """interpolate(im, x, y, out_size) (spatial_transformer.py:200-281), incl. im size != out size."""
    m = rl.st1()
    for name, n, ih, iw, oh, ow, c, seed in (('interp_same', 2, 48, 64, 48, 64, 3, 210),
                                             ('interp_resize_c1', 2, 40, 56, 48, 64, 1, 220)):
        im = synth.noise_image(n, ih, iw, c, seed)
        x = synth.uniform((n, oh, ow, 1), -1.2, 1.2, seed + 1)
        y = synth.uniform((n, oh, ow, 1), -1.2, 1.2, seed + 2)
        d_out = synth.randn((n, oh, ow, c), seed + 3)
        fx = dict(im=im, x=x, y=y, d_out=d_out)
        for tag, dt in (('ref', torch.float32), ('f64', torch.float64)):
            tf = rl.tf()
            tf.DEFAULT_FLOAT[0] = dt
            try:
                it, xt, yt = _tt(im, dt, True), _tt(x, dt, True), _tt(y, dt, True)
                out = m.interpolate(it, xt, yt, (oh, ow))
                (out * _tt(d_out, dt)).sum().backward()
                fx.update({tag + '_out': _np(out), tag + '_dim': _np(it.grad), tag + '_dx': _np(xt.grad),
                           tag + '_dy': _np(yt.grad)})
            finally:
                tf.DEFAULT_FLOAT[0] = torch.float32
        yield name, fx


def gen_loss_cases():
    """get_4_pts, feature_loss, img_loss (s_net_bundle_nobm.py) and temp_loss (train_bundle_nobm.py)."""
    n, h, w, gh, gw, mm = 3, 48, 64, 4, 4, 40
    head = synth.randn((n, 2 * (gh + 1) * (gw + 1)), 310, 0.12)       # network output; some vertices hit the clamp
    head2 = synth.randn((n, 2 * (gh + 1) * (gw + 1)), 311, 0.05)
    x1 = synth.smooth_image(n, h, w, 1, 312)
    x2 = synth.smooth_image(n, h, w, 1, 313)
    y1 = synth.smooth_image(n, h, w, 1, 314)
    y2 = synth.smooth_image(n, h, w, 1, 315)
    matches = synth.uniform((n, mm, 4), -1.1, 1.1, 316)
    mask = (synth.uniform((n, mm), 0, 1, 317) < 0.3).astype(np.float32)
    mask[2] = 0                                                        # max(sum(mask),1) branch
    flow = np.concatenate([synth.uniform((n, h, w, 1), -1.05, 1.05, 318), synth.uniform((n, h, w, 1), -1.05, 1.05, 319)], -1)
    fx = dict(head=head, head2=head2, x1=x1, x2=x2, y1=y1, y2=y2, matches=matches, mask=mask, flow=flow,
              grid=np.array([gh, gw]))
    for tag, dt in (('ref', torch.float32), ('f64', torch.float64)):
        tf = rl.tf()
        tf.DEFAULT_FLOAT[0] = dt
        try:
            ns = rl.s_net_namespace(h, w, gh, gw, n, mm)
            m = rl.st3(h, w, gh, gw)
            hd, hd2 = _tt(head, dt, True), _tt(head2, dt, True)
            pts1, pts2 = ns['get_4_pts'](hd, n)
            _, pts2b = ns['get_4_pts'](hd2, n)
            out1, black1, img1 = rl.quiet(m.transformer, _tt(x1, dt), pts2)
            out2, black2, img2 = rl.quiet(m.transformer, _tt(x2, dt), pts2b)
            f_loss, i_loss, warpped = rl.s_net_losses(ns, _tt(matches, dt), _tt(mask, dt), img1, out1, _tt(y1, dt), black1)
            ret1 = dict(output=out1, black_pix=black1.reshape(n, h, w, 1))
            ret2 = dict(output=out2, black_pix=black2.reshape(n, h, w, 1))
            t_loss = rl.train_temp_loss(h, w, n, ret1, ret2, _tt(flow, dt))
            fx.update({tag + '_pts1': _np(pts1), tag + '_pts2': _np(pts2), tag + '_pts2b': _np(pts2b),
                       tag + '_feature_loss': _np(f_loss), tag + '_img_loss': _np(i_loss), tag + '_temp_loss': _np(t_loss),
                       tag + '_stable_warpped': _np(warpped)})
            for nm, l in (('feature', f_loss), ('img', i_loss), ('temp', t_loss)):
                g = torch.autograd.grad(l, [hd, hd2], retain_graph=True, allow_unused=True)
                fx[tag + '_dhead_' + nm] = _np(g[0]) if g[0] is not None else np.zeros_like(head)
                fx[tag + '_dhead2_' + nm] = _np(g[1]) if g[1] is not None else np.zeros_like(head2)
        finally:
            tf.DEFAULT_FLOAT[0] = torch.float32
    yield 'losses', fx


def main(argv):
    assert rl.available(), 'needs /root/reference (build container only)'
    torch.set_num_threads(1)      # fixed summation order in the reference's reductions
    os.makedirs(OUT, exist_ok=True)
    want = set(argv)
    for gen in (gen_mesh_cases, gen_homography_cases, gen_interp_cases, gen_loss_cases):
        for name, fx in gen():
            if want and name not in want:
                continue
            path = os.path.join(OUT, name + '.npz')
            np.savez_compressed(path, **fx)
            print('%-26s %8.1f KB  %s' % (name, os.path.getsize(path) / 1024, ' '.join(sorted(fx))[:100]))


if __name__ == '__main__':
    main(sys.argv[1:])
