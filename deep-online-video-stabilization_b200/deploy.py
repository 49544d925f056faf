"""Deploy-side helpers with the reference's names (deploy_bundle.py).

    img_warped = warpRevBundle2(cv2.resize(frame, (width, height)), xmap, ymap)      # reference deploy_bundle.py:301

`warpRevBundle2` keeps the reference signature -- one uint8 frame [H,W,3] and the operator's x_map / y_map [H,W] -- and
accepts numpy arrays (as the reference's caller passes; the result comes back as a numpy array) or CUDA tensors (batched
[N,H,W,C] / [N,H,W] too; the result stays on the device).  The work is one C-ABI call, mgw_remap_bundle_u8.
"""
import numpy as np
import torch

from . import ops

__all__ = ['warpRevBundle2', 'warpRevBundle', 'warpRev', 'cvt_theta_mat_bundle', 'cvt_img2train', 'cv2_resize', 'StreamState', 'CropState']


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


def _pil_bilinear_windows(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (src/libImaging/Resample.c) for BILINEAR, in the same double
    arithmetic: per output sample the first source index, the window length and the 22-bit fixed-point weights."""
    import math
    scale = in_size / float(out_size)
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    first = np.zeros(out_size, np.int32)
    count = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [max(1.0 - abs((x + xmin - center + 0.5) * ss), 0.0) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            v = v / ww if ww != 0.0 else v
            kk[xx, x] = int(v * (1 << 22) + 0.5)
        first[xx], count[xx] = xmin, xmax
    return kk, first, count


def _cv_linear_table(n_out, n_in, horizontal):
    """resize.cpp: fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx), fx -= s; horizontally a tap outside the row is folded
    onto the border with weight 0, vertically the row indices are clamped and the weights kept; 11-bit weights cvRound(w*2048)"""
    f32 = np.float32
    scale = 1.0 / (n_out / float(n_in))
    tab = np.zeros((n_out, 4), np.int32)
    for d in range(n_out):
        f = f32((d + 0.5) * scale - 0.5)
        s0 = int(np.floor(f))
        f = f32(f - f32(s0))
        if horizontal:
            if s0 < 0:
                s0, f = 0, f32(0)
            if s0 >= n_in - 1:
                s0, f = n_in - 1, f32(0)
            i0, i1 = s0, min(s0 + 1, n_in - 1)
        else:
            i0, i1 = min(max(s0, 0), n_in - 1), min(max(s0 + 1, 0), n_in - 1)
        tab[d] = (i0, i1, int(np.rint(f32(f32(1) - f) * f32(2048))), int(np.rint(f * f32(2048))))
    return tab


_CVT_TABLES = {}
_RESIZE_TABLES = {}


def cv2_resize(img, dsize, device='cuda'):
    """cv2.resize(img, dsize) for a uint8 frame [H,W,C] (reference deploy_bundle.py:301 resizes the unstable colour frame to the
    network size before warpRevBundle2), byte-exact with OpenCV's INTER_LINEAR; dsize = (width, height) as in OpenCV.
    numpy in -> numpy out; a CUDA frame stays on the device."""
    as_numpy = isinstance(img, np.ndarray)
    im = (torch.as_tensor(np.ascontiguousarray(img)) if as_numpy else img).to(device=device if as_numpy else img.device, dtype=torch.uint8)
    H, W = int(im.shape[0]), int(im.shape[1])
    ow, oh = int(dsize[0]), int(dsize[1])
    key = (H, W, oh, ow, str(im.device))
    if key not in _RESIZE_TABLES:
        if W == 2 * ow and H == 2 * oh:
            _RESIZE_TABLES[key] = (None, None)
        else:
            _RESIZE_TABLES[key] = (torch.as_tensor(_cv_linear_table(ow, W, True)).to(im.device),
                                   torch.as_tensor(_cv_linear_table(oh, H, False)).to(im.device))
    dst = ops.resize_linear_u8(im.contiguous(), _RESIZE_TABLES[key][0], _RESIZE_TABLES[key][1], oh, ow)
    return dst.cpu().numpy() if as_numpy else dst


def cvt_img2train(img, crop_rate=1, height=288, width=512, device='cuda', as_numpy=False):
    """reference config.py:6-21 (`height`, `width` are its config globals): one BGR uint8 frame [H,W,3] -> the network's frame
    format, exact: cv2 BGR2GRAY, Pillow resize(BILINEAR) (to (width, height), or to the 1/crop_rate larger size followed by
    the centre crop), v * (1/255) - 0.5.  Returns a CUDA fp32 tensor [1,height,width,1]; as_numpy=True gives the reference's
    float64 numpy array (same values: the network's placeholder casts to fp32).  The coefficient tables are built once per
    shape on the host and kept on the device."""
    im = (torch.as_tensor(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img).to(device=device, dtype=torch.uint8)
    H, W = int(im.shape[0]), int(im.shape[1])
    key = (H, W, height, width, float(crop_rate), str(im.device))
    if key not in _CVT_TABLES:
        if crop_rate != 1:
            h = int(height / crop_rate); dh = int((h - height) / 2)
            w = int(width / crop_rate); dw = int((w - width) / 2)
        else:
            h, w, dh, dw = height, width, 0, 0
        kx, x0, xn = _pil_bilinear_windows(W, w)
        ky, y0, yn = _pil_bilinear_windows(H, h)
        tabs = (kx[dw:dw + width], x0[dw:dw + width], xn[dw:dw + width], ky[dh:dh + height], y0[dh:dh + height], yn[dh:dh + height])
        _CVT_TABLES[key] = tuple(torch.as_tensor(np.ascontiguousarray(t)).to(im.device) for t in tabs)
    out = ops.cvt_img2train_u8(im.contiguous(), _CVT_TABLES[key], height, width).reshape(1, height, width, 1)
    if not as_numpy:
        return out
    v = np.rint((out.cpu().numpy().astype(np.float64) + 0.5) * 255)          # the resampled byte, recovered exactly
    return v * (1. / 255) - 0.5


def cvt_theta_mat_bundle(Hs, height, width, grid_h, grid_w):
    """reference deploy_bundle.py:121-134 with its own numpy expressions (float64 on the host: 16 3x3 products), so that the
    pixel-space homographies the kernel receives are the reference's bit for bit; height / width / grid are its config globals."""
    from numpy.linalg import inv
    scale_mat = np.eye(3)
    scale_mat[0, 0] = width / 2.
    scale_mat[0, 2] = width / 2.
    scale_mat[1, 1] = height / 2.
    scale_mat[1, 2] = height / 2.
    Hs = np.asarray(Hs).reshape((grid_h, grid_w, 3, 3))
    return np.matmul(np.matmul(scale_mat, Hs), inv(scale_mat))


def warpRevBundle(img, Hs, grid=(4, 4), device=None):
    """reference deploy_bundle.py:148-173: one uint8 frame [H,W,3] and the operator's Hs ([gh,gw,9] or anything that reshapes to
    it, as `Hs[0]` of the reference's fetch) -> the frame with every mesh cell warped by its own homography
    (cv2.warpPerspective, WARP_INVERSE_MAP | INTER_LINEAR, byte-exact).  numpy in -> numpy out; a CUDA frame stays on the device."""
    as_numpy = isinstance(img, np.ndarray)
    dev = torch.device(device) if device is not None else (torch.device('cuda') if as_numpy else img.device)
    im = (torch.as_tensor(np.ascontiguousarray(img)) if as_numpy else img).to(device=dev, dtype=torch.uint8)
    h, w = im.shape[-3], im.shape[-2]
    gh, gw = int(grid[0]), int(grid[1])
    Hs_np = Hs.detach().cpu().numpy() if isinstance(Hs, torch.Tensor) else np.asarray(Hs)
    Hc = torch.as_tensor(np.ascontiguousarray(cvt_theta_mat_bundle(Hs_np, h, w, gh, gw), dtype=np.float64)).to(dev)
    dst = ops.warp_rev_bundle_u8(im.reshape(1, h, w, -1).contiguous(), Hc.reshape(1, gh, gw, 9), gh, gw)[0]
    return dst.cpu().numpy() if as_numpy else dst


def warpRev(img, theta, device=None):
    """reference deploy_bundle.py:100-118: cv2.warpPerspective(img, cvt_theta_mat(theta), WARP_INVERSE_MAP | INTER_LINEAR) for one
    3x3 homography in normalised coordinates -- the 1x1-grid case of warpRevBundle."""
    return warpRevBundle(img, np.asarray(theta.detach().cpu() if isinstance(theta, torch.Tensor) else theta).reshape(1, 1, 9),
                         grid=(1, 1), device=device)


class StreamState:
    """The per-video state of deploy_bundle.py on the device: the last `depth` stabilised frames and their black masks
    (the reference's before_frames / before_masks lists, :221-224) as rings, the 13-channel network input assembled from
    the index taps in one launch (:259-274), the refine re-feed (:292-295) and the per-frame update (:319-328).

        st = StreamState(first_frame)                    # first_frame: [H,W] fp32 (cvt_img2train output)
        for frame in video:
            in_x = st.assemble(frame)                    # [1,H,W,13] = 6 masks + 6 frames + current
            for _ in range(refine):
                img, black = net(in_x)                   # the network + the multi-grid warp
                st.refeed(in_x, img, black)              # in_x[..., -1] = img - black
            st.push(img, black)
    """

    def __init__(self, first_frame, depth=32, taps=(1, 2, 4, 8, 16, 32), use_masks=True, device='cuda', device_head=False):
        """device_head=True keeps the ring head in device memory: assemble / push then take no per-frame launch parameter, so the
        whole frame loop can be captured in one CUDA graph and replayed (the host-side `head` is not maintained in that mode)."""
        f = torch.as_tensor(first_frame, dtype=torch.float32).to(device)
        h, w = f.shape
        self.frames = f.reshape(1, h, w).repeat(depth, 1, 1).contiguous()
        self.masks = torch.zeros((depth, h, w), device=f.device, dtype=torch.float32)
        self.depth, self.taps, self.use_masks, self.head = depth, tuple(int(t) for t in taps), use_masks, depth - 1
        self.head_dev = torch.full((1,), depth - 1, device=f.device, dtype=torch.int32) if device_head else None

    def assemble(self, cur, out=None):
        cur = torch.as_tensor(cur, dtype=torch.float32).to(self.frames.device).reshape(self.frames.shape[1:]).contiguous()
        if self.head_dev is not None:
            return ops.stream_assemble_dev(self.frames, self.masks, self.head_dev, self.taps, cur, self.use_masks, out=out)
        return ops.stream_assemble(self.frames, self.masks, self.head, self.taps, cur, self.use_masks)

    def refeed(self, in_x, img, black):
        ops.stream_push(None, None, 0, img.reshape(self.frames.shape[1:]).contiguous(), black.reshape(self.frames.shape[1:]).contiguous(),
                        refeed_into=in_x)

    def push(self, img, black):
        img, black = img.reshape(self.frames.shape[1:]).contiguous(), black.reshape(self.frames.shape[1:]).contiguous()
        if self.head_dev is not None:
            ops.stream_push_dev(self.frames, self.masks, self.head_dev, img, black)
            return
        self.head = (self.head + 1) % self.depth
        ops.stream_push(self.frames, self.masks, self.head, img, black)

    def history(self):
        """(frames, masks) oldest first, as the reference's lists hold them"""
        head = int(self.head_dev.item()) if self.head_dev is not None else self.head
        order = [(head + 1 + k) % self.depth for k in range(self.depth)]
        return self.frames[order], self.masks[order]


class CropState:
    """The crop bookkeeping of deploy_bundle.py: `all_black` (:240) gathers every black mask the network produced (:291),
    and after the last frame the largest never-black rectangle on the 10-pixel corner lattice (:344-365) is cut out of
    every output frame (:368-370).

        crop = CropState(height, width)
        ... crop.add(black) after every network evaluation ...
        top, left, bottom, right = crop.rect()
        cut = crop.cut(frame)                            # frame[top:bottom+1, left:right+1, :]
    """

    def __init__(self, height, width, device='cuda'):
        self.all_black = torch.zeros((height, width), device=device, dtype=torch.int32)

    def add(self, black):
        ops.black_accumulate(self.all_black, black)

    def rect(self, step=10):
        r = [int(v) for v in ops.crop_rect(self.all_black, step).cpu()]
        if r[0] < 0:
            raise IndexError('no black-free rectangle: every lattice corner has been black (the reference fails on ans[3] here)')
        return r

    def cut(self, frame, step=10):
        t, l, b, r = self.rect(step)
        return frame[..., t:b + 1, l:r + 1, :]
