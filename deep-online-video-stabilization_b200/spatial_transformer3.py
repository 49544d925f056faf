"""Drop-in for the reference's spatial_transformer3.py (the multi-grid warp operator).

    from spatial_transformer3 import transformer          # reference s_net_bundle_nobm.py:16
    h_trans, black_pix, flow = transformer(x, pts2)        # reference s_net_bundle_nobm.py:307,332

Same call signature and return tuple as the reference (spatial_transformer3.py:19, :301): U [N,H,W,C] fp32 NHWC,
theta [N,gh+1,gw+1,2] absolute mesh vertices.  Where the reference reads height/width/grid_h/grid_w from the
global config (spatial_transformer3.py:16), this takes them from the tensor shapes.  The graph tensors
deploy_bundle.py:48-56 fetches by name are exposed through named_outputs() / NAMED.
"""
from . import functional as F
from .spatial_transformer import interpolate  # noqa: F401  (the reference file defines the same interpolate, :368-449)

NAMED = {}      # 'output_img', 'black_pix', 'Hs', 'x_map', 'y_map' of the most recent transformer() call


def transformer(U, theta, name='SpatialTransformer', **kwargs):
    """-> (output [N,H,W,C], black_pix [N,H,W], img [N,H,W,2]); differentiable w.r.t. U and theta."""
    if U.shape[0] == 0 and theta.shape[0] == 0:
        return F.empty_batch(U, theta, U.shape[1], U.shape[2])
    out, black, img, Hs = F.MeshWarp.apply(U, theta)
    NAMED.update(output_img=out, black_pix=black, Hs=Hs, x_map=img[..., 0:1], y_map=img[..., 1:2])
    return out, black, img


def named_outputs(U, theta):
    """the five tensors deploy_bundle.py:51-56 fetches: output_img, black_pix, get_Hs/Hs, x_map, y_map."""
    transformer(U, theta)
    return dict(NAMED)
