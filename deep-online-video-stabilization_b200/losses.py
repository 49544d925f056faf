"""Vertex builder and the loss epilogue fused onto the warped output (north_star part 4).

Mirrors the reference expressions; `batch_size` defaults to the leading dimension (the reference hard-wires the
config constant, s_net_bundle_nobm.py:348,352) and is the GLOBAL batch under data parallelism.
"""
from . import functional as F


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
