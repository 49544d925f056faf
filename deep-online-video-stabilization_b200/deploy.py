"""Deploy-side helpers with the reference's names (deploy_bundle.py).

    img_warped = warpRevBundle2(cv2.resize(frame, (width, height)), xmap, ymap)      # reference deploy_bundle.py:301

`warpRevBundle2` keeps the reference signature -- one uint8 frame [H,W,3] and the operator's x_map / y_map [H,W] -- and
accepts numpy arrays (as the reference's caller passes; the result comes back as a numpy array) or CUDA tensors (batched
[N,H,W,C] / [N,H,W] too; the result stays on the device).  The work is one C-ABI call, mgw_remap_bundle_u8.
"""
import numpy as np
import torch

from . import ops

__all__ = ['warpRevBundle2']


def warpRevBundle2(img, x_map, y_map, device=None):
    as_numpy = isinstance(img, np.ndarray)
    dev = torch.device(device) if device is not None else (torch.device('cuda') if as_numpy else img.device)

    def to_dev(a, dtype):
        t = torch.as_tensor(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        return t.to(device=dev, dtype=dtype)

    im, xm, ym = to_dev(img, torch.uint8), to_dev(x_map, torch.float32), to_dev(y_map, torch.float32)
    single = im.dim() == 3
    if single:
        im = im.unsqueeze(0)
    n, h, w, _ = im.shape
    xy = torch.stack([xm.reshape(n, h, w), ym.reshape(n, h, w)], dim=-1)
    dst = ops.remap_bundle_u8(im, xy)
    if single:
        dst = dst[0]
    return dst.cpu().numpy() if as_numpy else dst
