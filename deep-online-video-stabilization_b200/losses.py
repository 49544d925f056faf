"""Vertex builder and the loss epilogue fused onto the warped output (north_star part 4).

Mirrors the reference expressions; `batch_size` defaults to the leading dimension (the reference hard-wires the
config constant, s_net_bundle_nobm.py:348,352) and is the GLOBAL batch under data parallelism.
"""
import torch

from . import functional as F
from . import ops


def get_4_pts(theta, batch_size=None, grid=(4, 4), do_crop_rate=0.8):
    """reference s_net_bundle_nobm.py:29-71: network output [N,2*(gh+1)*(gw+1)] -> (pts1 [N,gh,gw,8], pts2 [N,gh+1,gw+1,2])."""
    return F.Vertices.apply(theta, int(grid[0]), int(grid[1]), float(do_crop_rate))


def img_loss(h_trans, y, black_pix, batch_size=None):
    """reference s_net_bundle_nobm.py:347-352."""
    n, h, w, _ = h_trans.shape
    return F.ImgLoss.apply(h_trans, y, black_pix.reshape(n, h, w), float(batch_size or n))


def transformer_img_loss(U, theta, y, batch_size=None):
    """transformer(U, theta) and img_loss(h_trans, y, black_pix) fused into the warp kernels (north_star part 4):
    -> (img_loss, h_trans, black_pix, flow).  Equivalent to reference s_net_bundle_nobm.py:332 + :347-352."""
    return F.MeshWarpImgLoss.apply(U, theta, y, float(batch_size or U.shape[0]))


def feature_loss(matches, mask, flow, batch_size=None):
    """reference s_net_bundle_nobm.py:335-343 -> (loss, stable_warpped [N,M,2])."""
    return F.FeatureLoss.apply(matches, mask, flow, float(batch_size or flow.shape[0]))


def temp_loss(out1, black1, out2, black2, flow, use_temp_loss=1.0, batch_size=None):
    """reference train_bundle_nobm.py:115-125."""
    n, h, w, _ = out1.shape
    return F.TempLoss.apply(out1, black1.reshape(n, h, w), out2, black2.reshape(n, h, w), flow, float(batch_size or n),
                            float(use_temp_loss))


# ---------------------------------------------------------------------------------------------------------------------
# vertex regularisers and the loss schedule (s_net_bundle_nobm.py:139-210,246-247,312-317,354-359; train_bundle_nobm.py:219-236)
def _grid_of(pts1=None, pts2=None):
    if pts1 is not None:
        return int(pts1.shape[1]), int(pts1.shape[2])
    return int(pts2.shape[1]) - 1, int(pts2.shape[2]) - 1


def vertex_losses(theta, pts1, pts2, do_crop_rate=0.8, batch_size=None):
    """All four head-only terms in one launch -> (id_loss_unscaled, black_pos_loss_ungated, distortion_loss, consistency_loss).
    id_loss_unscaled = mean|theta| (the reference multiplies by id_mul, :246); black_pos_loss_ungated = mean(black_err^2)
    (the reference multiplies by use_black_loss, :316).  batch_size: the divisor's batch (GLOBAL batch under data parallelism)."""
    gh, gw = _grid_of(pts1, pts2)
    n = float(batch_size or next(t for t in (theta, pts1, pts2) if t is not None).shape[0])
    sums = F.VertexLosses.apply(theta, pts1, pts2, gh, gw, float(do_crop_rate))
    terms = 2 * (max(gh - 1, 0) * (gw + 1) + (gh + 1) * max(gw - 1, 0))
    return (sums[0] / (n * 2 * (gh + 1) * (gw + 1)), sums[1] / (n * gh * gw * 8), sums[2] / (n * gh * gw * 2) / 8,
            sums[3] / (n * 2 * terms) if terms else sums[3] * 0)


def get_black_pos(pts, do_crop_rate=0.8):
    """reference s_net_bundle_nobm.py:139-148 -> black_err [N, gh*gw*8] (the display tensor ret['black_pos'] before squaring;
    the differentiable path is vertex_losses / get_black_pos_loss)."""
    gh, gw = _grid_of(pts)
    _, err = ops.vertex_losses_fwd(None, pts.detach().contiguous(), None, gh, gw, float(do_crop_rate), want_black_err=True)
    return err.reshape(pts.shape[0], -1)


def get_black_pos_loss(pts1, use_black_loss=1.0, do_crop_rate=0.8, batch_size=None):
    """reference s_net_bundle_nobm.py:312-317"""
    return vertex_losses(None, pts1, None, do_crop_rate, batch_size)[1] * use_black_loss


def get_distortion_loss(pts, batch_size=None):
    """reference s_net_bundle_nobm.py:168-184 (pts = pts1 [N,gh,gw,8])"""
    return vertex_losses(None, pts, None, 0.8, batch_size)[2]


def get_consistency_loss(pts, batch_size=None):
    """reference s_net_bundle_nobm.py:186-210 (pts = pts2 [N,gh+1,gw+1,2])"""
    return vertex_losses(None, None, pts, 0.8, batch_size)[3]


def loss_gates(i, no_theta_iter=1000000, do_temp_loss_iter=5000, do_theta_10_iter=-1, do_black_loss_iter=1000,
               do_theta_only_iter=100):
    """reference train_bundle_nobm.py:219-236 (defaults: configs/v2_93.py:39-43) -> dict(use_theta, use_temp, use_black, theta_only)
    for training iteration i."""
    use_theta = 0 if i > no_theta_iter else 1
    if i <= do_theta_10_iter:
        use_theta = 10
    return dict(use_theta=use_theta, use_temp=1 if i >= do_temp_loss_iter else 0, use_black=1 if i >= do_black_loss_iter else 0,
                theta_only=1 if i <= do_theta_only_iter else 0)


# configs/v2_93.py:7-13,44-48
V2_93_MULS = dict(feature_mul=1.0, theta_mul=400 / 2500, regu_mul=30 / 2500, img_mul=50.0, temp_mul=500.0, black_mul=300000 / 2500,
                  id_mul=10 / 2500, distortion_mul=1.0, consistency_mul=20.0, grid_theta_mul=0.0)


def total_loss(theta, pts1, pts2, img_loss, feature_loss, regu_loss=0.0, use_black_loss=1.0, use_theta_only=0.0, mul=None,
               do_crop_rate=0.8, batch_size=None):
    """reference s_net_bundle_nobm.py:308-317,354-359 for one pass: the head-only terms come from ONE launch of the vertex-loss
    kernel; img_loss / feature_loss are the (already computed) image-side terms, regu_loss the network's weight regulariser.
    -> (total_loss, dict of the weighted parts as in the reference's `ret`)."""
    m = dict(V2_93_MULS)
    m.update(mul or {})
    idl, black, dist, cons = vertex_losses(theta, pts1, pts2, do_crop_rate, batch_size)
    theta_loss = idl * m['id_mul']                      # id_loss == id2_loss (:246-247)
    black = black * use_black_loss
    total = theta_loss * m['theta_mul'] + theta_loss * m['grid_theta_mul'] + (1 - use_theta_only) * (
        img_loss * m['img_mul'] + regu_loss * m['regu_mul'] + black * m['black_mul'] + dist * m['distortion_mul'] +
        cons * m['consistency_mul'] + feature_loss * m['feature_mul'])
    parts = dict(theta_loss=theta_loss * m['theta_mul'], grid_theta_loss=theta_loss * m['grid_theta_mul'], black_loss=black * m['black_mul'],
                 distortion_loss=dist * m['distortion_mul'], consistency_loss=cons * m['consistency_mul'],
                 feature_loss=feature_loss * m['feature_mul'], img_loss=img_loss * m['img_mul'],
                 regu_loss=torch.as_tensor(regu_loss) * m['regu_mul'])
    return total, parts


# the weighted parts of one pass, in the order mgw_train_pass_fwd reports them (result[1:]), keys as in the reference's `ret`
PASS_PARTS = ('theta_loss', 'grid_theta_loss', 'black_loss', 'distortion_loss', 'consistency_loss', 'feature_loss', 'img_loss',
              'regu_loss')


def pass_coef(n, gh, gw, use_black_loss=1.0, use_theta_only=0.0, mul=None):
    """The 11 host coefficients of mgw_train_pass_fwd/bwd for a (global) batch n: reference s_net_bundle_nobm.py:308-317,354-359
    with the element counts of the four means folded in."""
    m = dict(V2_93_MULS)
    m.update(mul or {})
    n = float(n)
    terms = 2 * (max(gh - 1, 0) * (gw + 1) + (gh + 1) * max(gw - 1, 0))
    idw = m['theta_mul'] + m['grid_theta_mul']
    return (m['id_mul'] * idw / (n * 2 * (gh + 1) * (gw + 1)),
            float(use_black_loss) * m['black_mul'] / (n * gh * gw * 8),
            m['distortion_mul'] / (n * gh * gw * 2) / 8,
            m['consistency_mul'] / (n * 2 * terms) if terms else 0.0,
            m['img_mul'], m['feature_mul'], m['regu_mul'],
            m['theta_mul'] / idw if idw else 0.0, m['grid_theta_mul'] / idw if idw else 0.0,
            1.0 / n, 1.0 - float(use_theta_only))


def train_pass(theta, x, y, matches, mask, regu_loss=None, use_black_loss=1.0, use_theta_only=0.0, mul=None, grid=(4, 4),
               do_crop_rate=0.8, batch_size=None):
    """Everything of one pass after the network head in ONE autograd node (7 + 5 launches): get_4_pts, transformer + img_loss,
    feature_loss, the vertex regularisers and the weighted total (reference s_net_bundle_nobm.py:303-359).  theta = the head's
    output [N, 2(gh+1)(gw+1)], x = the frame to warp.  -> (total_loss, parts dict as in the reference's `ret`, h_trans, black_pix,
    flow, stable_warpped, pts2).  h_trans may feed another differentiable consumer (temp_loss); flow and pts2 are for display
    (take them from transformer() / get_4_pts() if they need a gradient of their own)."""
    gh, gw = int(grid[0]), int(grid[1])
    coef = pass_coef(batch_size or x.shape[0], gh, gw, use_black_loss, use_theta_only, mul)
    regu = regu_loss if isinstance(regu_loss, torch.Tensor) else None
    if regu is None and regu_loss:
        regu = torch.full((1,), float(regu_loss), device=x.device)
    total, parts, out, black, flow, warpped, pts2 = F.TrainPass.apply(theta, x, y, matches, mask, regu, coef, gh, gw, float(do_crop_rate))
    return total, dict(zip(PASS_PARTS, parts.unbind(0))), out, black, flow, warpped, pts2
