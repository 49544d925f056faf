"""TEST INFRASTRUCTURE -- not product code.

Closed-form NumPy restatement of the vertex regularisers of s_net_bundle_nobm.py (values and gradients), checked against
the reference's own functions run on the shim (tests/golden/vertex_losses.npz, oracle/make_golden.py gen_vertex_loss_cases).
  id_loss          :246-247    mean |theta|  (times id_mul)
  black_pos_loss   :139-148,312-317
  distortion_loss  :150-184
  consistency_loss :186-210
  loss_gates       train_bundle_nobm.py:219-236
  total_loss       s_net_bundle_nobm.py:354-359, train_bundle_nobm.py:141
"""
import numpy as np

# (p0, p1, p2, clock, hw) of the eight calc_distortion_loss calls, :174-181
EDGES = ((0, 1, 3, 0, 0), (1, 3, 2, 0, 1), (3, 2, 0, 0, 0), (2, 0, 1, 0, 1),
         (1, 0, 2, 1, 0), (0, 2, 3, 1, 1), (2, 3, 1, 1, 0), (3, 1, 0, 1, 1))


def id_loss(theta):
    t = np.asarray(theta)
    return np.abs(t).mean(dtype=t.dtype), np.sign(t) / t.dtype.type(t.size)


def black_pos_loss(pts1, do_crop_rate=0.8, use_black_loss=1.0):
    p = np.asarray(pts1)
    one = p.dtype.type(1.0) / p.dtype.type(do_crop_rate)
    err = np.where(p > one, p - one, 0) + np.where(-one > p, -one - p, 0)
    g = 2 * err * np.where(p > one, 1, np.where(-one > p, -1, 0)) * use_black_loss / p.size
    return (err * err * use_black_loss).mean(dtype=p.dtype), g.astype(p.dtype), err


def distortion_loss(pts1, gh, gw):
    p = np.asarray(pts1)
    dt = p.dtype.type
    c = p.reshape(-1, 2, 4)                                   # [cells, (x|y), corner]
    h, w = 2.0 / gh, 2.0 / gw
    total = np.zeros((), p.dtype)
    g = np.zeros_like(c)
    for a, b, cc, clock, hw in EDGES:
        k = dt(np.float32(h / w if hw == 0 else w / h))          # R is a float32 constant in the reference (:161)
        R = np.array([[0, k], [-k, 0]] if clock else [[0, -k], [k, 0]], p.dtype)
        v = c[:, :, b] - c[:, :, a]
        e = v @ R.T - (c[:, :, cc] - c[:, :, b])              # [cells, 2]
        total = total + (e * e).sum(dtype=p.dtype)
        u = 2 * e
        t = u @ R                                             # R^T u per row
        g[:, :, a] -= t
        g[:, :, b] += t + u
        g[:, :, cc] -= u
    cnt = dt(c.shape[0] * 2 * 8)
    return total / cnt, (g / cnt).reshape(p.shape)


def consistency_loss(pts2, gh, gw):
    p = np.asarray(pts2)
    n = p.shape[0]
    terms = 2 * (max(gh - 1, 0) * (gw + 1) + (gh + 1) * max(gw - 1, 0))      # len(errs), :194-203
    g = np.zeros_like(p)
    if terms == 0:
        return np.zeros((), p.dtype), g
    total = np.zeros((), p.dtype)
    for axis, size in ((1, gh), (2, gw)):
        q = np.moveaxis(p, axis, 1)
        gq = np.moveaxis(g, axis, 1)
        for r in range(1, size):                              # triple (r-1, r, r+1), listed under both of its ends
            d = 2 * q[:, r] - q[:, r - 1] - q[:, r + 1]
            total = total + 2 * (d * d).sum(dtype=p.dtype)
            gq[:, r] += 2 * 2 * 2 * d
            gq[:, r - 1] -= 2 * 2 * d
            gq[:, r + 1] -= 2 * 2 * d
    cnt = p.dtype.type(n * 2 * terms)
    return total / cnt, g / cnt


def loss_gates(i, no_theta_iter=1000000, do_temp_loss_iter=5000, do_theta_10_iter=-1, do_black_loss_iter=1000,
               do_theta_only_iter=100):
    """train_bundle_nobm.py:219-236 -> (use_theta, use_temp, use_black, theta_only)"""
    use_theta = 0 if i > no_theta_iter else 1
    use_temp = 1 if i >= do_temp_loss_iter else 0
    if i <= do_theta_10_iter:
        use_theta = 10
    use_black = 1 if i >= do_black_loss_iter else 0
    theta_only = 1 if i <= do_theta_only_iter else 0
    return use_theta, use_temp, use_black, theta_only


def total_loss(theta_loss, grid_theta_loss, img_loss, regu_loss, black_pos_loss, distortion_loss, consistency_loss, feature_loss,
               use_theta_only, mul):
    """s_net_bundle_nobm.py:356-358 for one of the two passes"""
    return theta_loss * mul['theta_mul'] + grid_theta_loss * mul['grid_theta_mul'] + ((1 - use_theta_only) * (
        img_loss * mul['img_mul'] + regu_loss * mul['regu_mul'] + black_pos_loss * mul['black_mul'] +
        distortion_loss * mul['distortion_mul'] + consistency_loss * mul['consistency_mul'] + feature_loss * mul['feature_mul']))
