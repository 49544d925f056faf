"""Drop-in for the reference's spatial_transformer.py.

    from spatial_transformer import *                      # reference train_bundle_nobm.py:16
    out = interpolate(im, x_flow, y_flow, (height, width)) # reference train_bundle_nobm.py:117-118,134
    out, black = transformer(U, theta9, (h, w))            # reference spatial_transformer.py:18,193

`transformer` here is the single-homography operator (theta [N,9] or [N,3,3], divided by theta[2,2]); passing a
mesh [N,gh+1,gw+1,2] dispatches to the multi-grid operator so that the signature BASELINE.json's north_star
names -- transformer(U, theta, out_size) -- covers both.
"""
from . import functional as F

__all__ = ['transformer', 'interpolate']


def transformer(U, theta, out_size, name='SpatialTransformer', **kwargs):
    n = U.shape[0]
    if n == 0 and theta.shape[0] == 0:
        out, black, img = F.empty_batch(U, theta, int(out_size[0]), int(out_size[1]))
        return (out, black, img) if (theta.dim() == 4 and theta.shape[-1] == 2) else (out, U.new_zeros((0, U.shape[1], U.shape[2])))
    if theta.dim() == 4 and theta.shape[-1] == 2:
        if tuple(int(v) for v in out_size) != (U.shape[1], U.shape[2]):
            raise ValueError('the multi-grid warp produces an output of the input size (reference spatial_transformer3.py:289)')
        out, black, img, _ = F.MeshWarp.apply(U, theta)
        return out, black, img
    if theta.numel() != n * 9:
        raise ValueError('theta must be [N,9] / [N,3,3] (homography) or [N,gh+1,gw+1,2] (mesh)')
    return F.HomographyWarp.apply(U, theta.reshape(n, 9), (int(out_size[0]), int(out_size[1])))


def interpolate(im, x, y, out_size, name='SpatialInterpolate', **kwargs):
    """im [N,IH,IW,C]; x, y [N,out_h,out_w,1] normalised coordinates -> [N,out_h,out_w,C]."""
    if im.shape[0] == 0 and x.shape[0] == 0:
        return F.empty_batch(im, x, int(out_size[0]), int(out_size[1]))[0] + y.sum() * 0
    return F.Interpolate.apply(im, x, y, (int(out_size[0]), int(out_size[1])))
