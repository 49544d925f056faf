"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Differentiable torch-CPU restatement ("port") of the reference's multi-grid warp path, vectorised
over cells/pixels but keeping the reference's fp32 operation order.  Runs in fp32 (reference
behaviour) or fp64 (arbiter) depending on the dtype of its inputs.  Uses:

  * gradient oracle: torch autograd cuts the graph at floor / int cast / clip exactly where TF does
    (SURVEY.md 8a-bwd), so .backward() through these functions is the reference's backward;
  * bench.py's cpu_baseline / `--impl reference` arm (kind "port"; TensorFlow does not exist here).

Pinned against tests/golden/*.npz (outputs of the unmodified reference sources on the tensorflow
shim) by tests/test_oracle_golden.py.  The bit-exact forward checker is oracle/mgw_oracle.c; this
file emulates fmaf through fp64 and can differ from it by one ulp on a ~1e-9 fraction of pixels.

Citations are file:line under /root/reference.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference arm may import this module.
"""
import torch


def _fma(a, b, c):
    """fp32 fused multiply-add emulated in fp64 (exact product, one extra rounding); plain in fp64."""
    if a.dtype == torch.float64:
        return a * b + c
    return (a.double() * b.double() + c.double()).float()


def get_4_pts(head, gh, gw, do_crop_rate=0.8):
    """s_net_bundle_nobm.py:29-71: head [N,2*(gh+1)*(gw+1)] -> (pts1 [N,gh,gw,8], pts2 [N,gh+1,gw+1,2])."""
    n, dt = head.shape[0], head.dtype
    h, w = 2.0 / gh, 2.0 / gw
    base = torch.tensor([[[j * w - 1, i * h - 1] for j in range(gw + 1)] for i in range(gh + 1)],
                        dtype=torch.float32).to(dt)                                                             # :44-46 (fp32 constants)
    lim = (torch.ones((), dtype=dt) / torch.tensor(do_crop_rate, dtype=dt))                                    # :37
    p = base[None] + head.reshape(n, gh + 1, gw + 1, 2)                                                        # :47,:55
    pts2 = torch.minimum(torch.maximum(p, -1 * lim), lim)                                                      # :58
    tl, tr, bl, br = pts2[:, :-1, :-1], pts2[:, :-1, 1:], pts2[:, 1:, :-1], pts2[:, 1:, 1:]
    g = torch.stack([tl, tr, bl, br], dim=-1)                      # [N,gh,gw,2,4]   (:65)
    return g.reshape(n, gh, gw, 8), pts2                           # (:66-69)


def solve_h(theta):
    """get_Hs/get_H/pinv, spatial_transformer3.py:144-198: theta [N,gh+1,gw+1,2] -> Hs [N,gh,gw,9]."""
    n, gh1, gw1, _ = theta.shape
    gh, gw, dt = gh1 - 1, gw1 - 1, theta.dtype
    h, w = 2.0 / gh, 2.0 / gw
    ori = torch.tensor([[[j * w - 1, i * h - 1, j * w - 1 + w, i * h - 1, j * w - 1, i * h - 1 + h, j * w - 1 + w, i * h - 1 + h]
                         for j in range(gw)] for i in range(gh)], dtype=torch.float32).to(dt)                  # :182-189 (fp32 constants)
    x, y = ori[..., 0::2].expand(n, gh, gw, 4), ori[..., 1::2].expand(n, gh, gw, 4)                            # :154-155
    tar = torch.stack([theta[:, :-1, :-1], theta[:, :-1, 1:], theta[:, 1:, :-1], theta[:, 1:, 1:]], dim=3)    # :191-193
    u, v = tar[..., 0], tar[..., 1]                                                                            # :156-157
    one, zero = torch.ones_like(x), torch.zeros_like(x)
    top = torch.stack([x, y, one, zero, zero, zero, -x * u, -y * u], dim=-1)                                   # :160-163
    bot = torch.stack([zero, zero, zero, x, y, one, -x * v, -y * v], dim=-1)                                   # :164-167
    A = torch.cat([top, bot], dim=3)                               # [N,gh,gw,8,8]
    b = torch.cat([u, v], dim=3).unsqueeze(-1)                     # :169-170
    eps = torch.eye(8, dtype=dt) * torch.tensor(1e-4, dtype=torch.float32).to(dt) if dt == torch.float32 \
        else torch.eye(8, dtype=dt) * 1e-4                                                                     # :145
    hh = torch.matmul(torch.linalg.inv(A + eps), b).squeeze(-1)                                                # :173
    return torch.cat([hh, torch.ones(n, gh, gw, 1, dtype=dt)], dim=-1)


def _linspace(num, dt):
    """tf.linspace(-1,1,num): start + i*step (spatial_transformer3.py:205-206)."""
    start = torch.tensor(-1.0, dtype=dt)
    step = (torch.tensor(1.0, dtype=dt) - start) / torch.tensor(num - 1, dtype=dt)
    return start + step * torch.arange(num, dtype=dt)


def project(Hpix, height, width):
    """:248-260. Hpix broadcastable to [N,height,width,9]; returns xn, yn [N,height,width]."""
    dt = Hpix.dtype
    X = _linspace(width, dt)[None, None, :]
    Y = _linspace(height, dt)[None, :, None]
    hk = [Hpix[..., k] for k in range(9)]

    def row(a, b, c):
        t, by = torch.broadcast_tensors(a * X, b * 0 + Y)
        t = _fma(b.expand_as(t), by, t)
        return t + c
    xs, ys, zs = row(hk[0], hk[1], hk[2]), row(hk[3], hk[4], hk[5]), row(hk[6], hk[7], hk[8])
    sign = torch.where(zs >= 0, torch.ones_like(zs), torch.zeros_like(zs)) * 2 - 1                             # :257
    eps = torch.tensor(1e-8, dtype=torch.float32).to(dt) if dt == torch.float32 else 1e-8
    zs = zs + sign * eps                                                                                       # :258
    return xs / zs, ys / zs


def black_of(xn, yn):
    """:282-286"""
    cond = (-1 > xn) | (xn > 1) | (-1 > yn) | (yn > 1)
    return torch.where(cond, torch.ones_like(xn), torch.zeros_like(xn))


def cell_index(height, width, gh, gw):
    """:227-243: i = min(r // floor(H/gh), gh-1); last cell absorbs the remainder."""
    ci = torch.clamp(torch.arange(height) // (height // gh), max=gh - 1)
    cj = torch.clamp(torch.arange(width) // (width // gw), max=gw - 1)
    return ci, cj


def interpolate_core(im, xn, yn):
    """_interpolate, spatial_transformer3.py:62-123. im [N,IH,IW,C]; xn,yn [N,...] -> [N,...,C]."""
    n, ih, iw, c = im.shape
    x = (xn + 1.0) * float(iw) / 2.0                                                                           # :81-82
    y = (yn + 1.0) * float(ih) / 2.0
    big = 2147483648.0

    def fl(v):          # tf.cast(tf.floor(v),'int32') with x86 out-of-range -> INT_MIN
        f = torch.floor(v.detach())
        ok = f.abs() < big
        return torch.where(ok, f, torch.full_like(f, -big)).to(torch.int64)
    x0, y0 = fl(x), fl(y)
    x1, y1 = x0 + 1, y0 + 1
    x0, x1 = x0.clamp(0, iw - 1), x1.clamp(0, iw - 1)                                                          # :90-93
    y0, y1 = y0.clamp(0, ih - 1), y1.clamp(0, ih - 1)
    shape = [n] + [1] * (xn.dim() - 1)
    base = (torch.arange(n) * (ih * iw)).reshape(shape)                                                        # :96
    flat = im.reshape(-1, c)
    Ia, Ib = flat[base + y0 * iw + x0], flat[base + y1 * iw + x0]                                              # :97-111
    Ic, Id = flat[base + y0 * iw + x1], flat[base + y1 * iw + x1]
    x0f, x1f, y0f, y1f = x0.to(im.dtype), x1.to(im.dtype), y0.to(im.dtype), y1.to(im.dtype)
    wa = ((x1f - x) * (y1f - y)).unsqueeze(-1)                                                                 # :118-121
    wb = ((x1f - x) * (y - y0f)).unsqueeze(-1)
    wc = ((x - x0f) * (y1f - y)).unsqueeze(-1)
    wd = ((x - x0f) * (y - y0f)).unsqueeze(-1)
    return ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id                                                           # :122


def warp_given_h(U, Hs):
    """_transform3 after get_Hs (:227-301) -> (output_img, black_pix, img[N,H,W,2])."""
    n, height, width, _ = U.shape
    gh, gw = Hs.shape[1:3]
    ci, cj = cell_index(height, width, gh, gw)
    Hpix = Hs[:, ci][:, :, cj]                                     # [N,H,W,9]
    xn, yn = project(Hpix, height, width)
    out = interpolate_core(U, xn, yn)
    return out, black_of(xn, yn), torch.stack([xn, yn], dim=-1)


def transformer(U, theta):
    """spatial_transformer3.transformer(U, theta) (:19,:301): mesh theta [N,gh+1,gw+1,2]."""
    Hs = solve_h(theta)
    out, black, img = warp_given_h(U, Hs)
    return out, black, img, Hs


def transformer_homography(U, theta, out_size):
    """spatial_transformer.transformer(U, theta[N,9], out_size) (spatial_transformer.py:143-193)."""
    n = U.shape[0]
    th = theta.reshape(n, 9)
    th = th / th[:, 8:9]                                                                                       # :151-153
    xn, yn = project(th[:, None, None, :], out_size[0], out_size[1])
    return interpolate_core(U, xn, yn), black_of(xn, yn)


def interpolate(im, x, y, out_size):
    """spatial_transformer.interpolate (spatial_transformer.py:200-281)."""
    n, c = im.shape[0], im.shape[3]
    out = interpolate_core(im, x.reshape(n, -1), y.reshape(n, -1))
    return out.reshape(n, out_size[0], out_size[1], c)


def warp_pts(pts, flow):
    """s_net_bundle_nobm.py:215-230 (tf.round = half-to-even)."""
    n, height, width, _ = flow.shape
    x = torch.round(torch.clamp((pts[:, :, 0] + 1) / 2 * width, 0, width - 1)).to(torch.int64)
    y = torch.round(torch.clamp((pts[:, :, 1] + 1) / 2 * height, 0, height - 1)).to(torch.int64)
    fl = flow.reshape(n, height * width, 2)
    return torch.gather(fl, 1, (x + y * width).unsqueeze(-1).expand(-1, -1, 2))


def feature_loss(matches, mask, flow):
    """s_net_bundle_nobm.py:335-343."""
    warpped = warp_pts(matches[:, :, :2], flow)
    before = (warpped - matches[:, :, 2:]).abs().sum(2)
    after = (before * mask).sum(1) / torch.clamp(mask.sum(1), min=1)
    return after.mean(), warpped


def img_loss(h_trans, y, black_pix):
    """s_net_bundle_nobm.py:347-352 (batch_size = leading dim)."""
    n = h_trans.shape[0]
    nb = 1 - black_pix.reshape(n, h_trans.shape[1], h_trans.shape[2], 1)
    err = (h_trans - y) * nb
    return ((err * err).sum((1, 2, 3)) / (nb.sum((1, 2, 3)) + 1e-8)).sum(0) / n


def temp_loss(out1, black1, out2, black2, flow, use_temp_loss=1.0):
    """train_bundle_nobm.py:115-125."""
    n, height, width, _ = out1.shape
    xf, yf = flow[..., 0:1], flow[..., 1:2]
    o2 = interpolate(out2, xf, yf, (height, width))
    nb2 = interpolate(1 - black2.reshape(n, height, width, 1), xf, yf, (height, width))
    noblack = (1 - black1.reshape(n, height, width, 1)) * nb2
    err = (out1 - o2) * noblack
    return ((err * err).sum((1, 2, 3)) / (noblack.sum((1, 2, 3)) + 1e-8)).sum(0) / n * use_temp_loss
