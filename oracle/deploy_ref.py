"""TEST INFRASTRUCTURE -- numpy restatement of the deploy-side colour-frame warp, deploy_bundle.py:136-146
(`warpRevBundle2`): the network's x_map / y_map are smoothed by a /4 then x4 cv2.resize, turned into pixel coordinates
and used to cv2.remap the uint8 frame.  The arithmetic lives in a third-party dependency that is not under
/root/reference (OpenCV, version not pinned by the reference; 4.13.0 in this container); restated here from its
published algorithm and pinned against cv2 itself (tests/golden/deploy_remap.npz, oracle/make_golden.py):

  cv2.resize(float32, INTER_LINEAR)   modules/imgproc/src/resize.cpp, resizeGeneric_ + HResizeLinear / VResizeLinear:
      fx = (float)((dx + 0.5) * scale - 0.5), sx = floor(fx), fx -= sx; horizontally a tap that falls outside is
      folded onto the border with weight 0 (sx < 0 -> sx = 0, fx = 0; sx >= w-1 -> sx = w-1, fx = 0); vertically the ROW
      INDICES are clamped and the weights kept.  D = S[sx]*(1-fx) + S[sx+1]*fx per row, then rows*(1-fy) + rows'*fy, all in
      fp32 with separate multiplies and adds.  That is OpenCV's plain C++ path (cv2.setUseOptimized(False)): its
      SIMD-dispatched path (FMA) differs from it by 1 ulp on about a third of the map values, i.e. the reference's own
      result depends on the OpenCV build; parity is pinned to the plain path and the effect of the other one on the
      uint8 output is recorded in the fixture (fraction of differing bytes, max difference).
  cv2.remap(uint8, float32 maps, INTER_LINEAR, BORDER_CONSTANT 0)   modules/imgproc/src/imgwarp.cpp, remapBilinear:
      sx = cvRound(x*32), sy = cvRound(y*32) (round half to even), taps at (sx>>5, sy>>5) clamped to int16, 5-bit
      fractions, integer weights (32-fx)(32-fy)*32 ... summing to 2^15, out = (sum + 2^14) >> 15; taps outside the image
      read 0.  Identical in both OpenCV paths.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

f32 = np.float32


def _scale(n_out, n_in):
    """resize.cpp: inv_scale = dsize / (double)ssize; scale = 1. / inv_scale"""
    return 1.0 / (n_out / float(n_in))


def _hcoef(n_out, n_in):
    scale = _scale(n_out, n_in)
    i0 = np.zeros(n_out, np.int64)
    a = np.zeros((n_out, 2), np.float32)
    for d in range(n_out):
        fx = f32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = f32(fx - f32(sx))
        if sx < 0:
            sx, fx = 0, f32(0)
        if sx >= n_in - 1:
            sx, fx = n_in - 1, f32(0)
        i0[d] = sx
        a[d] = (f32(1) - fx, fx)
    return i0, np.minimum(i0 + 1, n_in - 1), a


def _vcoef(n_out, n_in):
    scale = _scale(n_out, n_in)
    i0 = np.zeros(n_out, np.int64)
    i1 = np.zeros(n_out, np.int64)
    a = np.zeros((n_out, 2), np.float32)
    for d in range(n_out):
        fy = f32((d + 0.5) * scale - 0.5)
        sy = int(np.floor(fy))
        fy = f32(fy - f32(sy))
        i0[d] = min(max(sy, 0), n_in - 1)
        i1[d] = min(max(sy + 1, 0), n_in - 1)
        a[d] = (f32(1) - fy, fy)
    return i0, i1, a


def resize_linear(s, out_w, out_h):
    """cv2.resize(s, (out_w, out_h)) for a float32 [H,W] array, plain C++ path."""
    s = np.asarray(s, np.float32)
    sh, sw = s.shape
    ix0, ix1, ax = _hcoef(out_w, sw)
    iy0, iy1, ay = _vcoef(out_h, sh)
    h = s[:, ix0] * ax[:, 0] + s[:, ix1] * ax[:, 1]
    return (h[iy0, :] * ay[:, 0:1] + h[iy1, :] * ay[:, 1:2]).astype(np.float32)


def remap_linear_u8(img, x, y):
    """cv2.remap(img, x, y, cv2.INTER_LINEAR) for uint8 [H,W,C] and float32 maps [OH,OW]; constant border 0."""
    H, W = img.shape[:2]
    def cv_round(v):            # cvtss2si: round half to even; out of range / NaN -> INT_MIN
        r = np.rint(v)
        bad = ~(np.abs(r) < 2147483648.0)
        return np.where(bad, -2147483648.0, r).astype(np.int64)

    sx = cv_round(np.asarray(x, np.float32) * f32(32))
    sy = cv_round(np.asarray(y, np.float32) * f32(32))
    ix = np.clip(sx >> 5, -32768, 32767)
    iy = np.clip(sy >> 5, -32768, 32767)
    fx, fy = sx & 31, sy & 31
    w = [(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32]

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        return v * ok[..., None]

    s = (tap(iy, ix) * w[0][..., None] + tap(iy, ix + 1) * w[1][..., None] + tap(iy + 1, ix) * w[2][..., None] +
         tap(iy + 1, ix + 1) * w[3][..., None])
    return np.clip((s + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def smooth_maps(x_map, y_map, rate=4):
    """deploy_bundle.py:139-143: /rate then x rate resize, then to pixel coordinates ((m + 1) / 2 * size in fp32)."""
    H, W = x_map.shape
    xs = resize_linear(resize_linear(x_map, int(W / rate), int(H / rate)), W, H)
    ys = resize_linear(resize_linear(y_map, int(W / rate), int(H / rate)), W, H)
    return ((xs + f32(1)) / f32(2) * f32(W)).astype(np.float32), ((ys + f32(1)) / f32(2) * f32(H)).astype(np.float32)


def warp_rev_bundle2(img, x_map, y_map):
    """deploy_bundle.py:136-146 for one uint8 frame [H,W,3] and the network's x_map / y_map [H,W]."""
    xp, yp = smooth_maps(np.asarray(x_map, np.float32), np.asarray(y_map, np.float32))
    return remap_linear_u8(img, xp, yp)


# ---------------------------------------------------------------------------------------------------------------------
# streaming state of deploy_bundle.py (:204-232 initialisation, :259-274 input assembly, :284-295 refine, :319-328 update)
class StreamStateRef:
    """Restatement of the reference's per-frame list handling: before_frames / before_masks are lists of `depth` arrays
    [1,H,W,1]; the network input is [masks[-i] for i in taps] + [frames[-i] for i in taps] + [current frame] along the
    channel axis; after the network has run, frame = img + black*(-1) is appended (and the oldest entry dropped)."""

    def __init__(self, first_frame, depth=32, taps=(1, 2, 4, 8, 16, 32), use_masks=True):
        h, w = first_frame.shape
        f = np.asarray(first_frame, np.float32).reshape(1, h, w, 1)
        self.frames = [f.copy() for _ in range(depth)]                                  # :221-223
        self.masks = [np.zeros((1, h, w, 1), np.float32) for _ in range(depth)]         # :224
        self.taps, self.use_masks, self.h, self.w = tuple(taps), use_masks, h, w

    def assemble(self, cur):
        in_x = []
        if self.use_masks:
            in_x += [self.masks[-i] for i in self.taps]
        in_x += [self.frames[-i] for i in self.taps]
        in_x.append(np.asarray(cur, np.float32).reshape(1, self.h, self.w, 1))
        return np.concatenate(in_x, axis=3)

    def frame_of(self, img, black):
        return (np.asarray(img, np.float32).reshape(self.h, self.w) + np.asarray(black, np.float32).reshape(self.h, self.w) *
                np.float32(-1)).astype(np.float32)

    def push(self, img, black):
        self.frames.append(self.frame_of(img, black).reshape(1, self.h, self.w, 1))
        self.masks.append(np.asarray(black, np.float32).reshape(1, self.h, self.w, 1))
        self.frames.pop(0)
        self.masks.pop(0)


def stream_fake_net(in_x, k):
    """deterministic stand-in for the network (sess.run at deploy_bundle.py:286) used by the stream fixtures and tests:
    output image and black mask [1,H,W,1] / [1,H,W] from the assembled input [1,H,W,13] and the frame number"""
    h, w = in_x.shape[1:3]
    r = np.random.RandomState(1000 + k)
    img = (0.7 * in_x[0, :, :, -1] + 0.3 * in_x[0, :, :, 6] + 0.05 * r.standard_normal((h, w))).astype(np.float32)
    black = (r.random_sample((h, w)) < 0.1).astype(np.float32)
    return img.reshape(1, h, w, 1), black.reshape(1, h, w)


# ---------------------------------------------------------------------------------------------------------------------
# crop of deploy_bundle.py (:240,291 accumulation; :344-365 search)
def black_accumulate(all_black, black):
    """:291  all_black = all_black + np.round(black).astype(np.int64)"""
    return all_black + np.round(np.asarray(black, np.float32).reshape(all_black.shape)).astype(np.int64)


def crop_rect(all_black, step=10):
    """:344-365 restated without the per-pixel loops: for a corner (i, j) on the lattice the widest black-free rectangle
    of height hh-i+1 ends at j + min(run[i..hh][j]) - 1, run[r][c] = count of consecutive zeros from (r, c) rightwards; the
    answer is the FIRST (i, j, hh) in loop order reaching the maximum area (the reference updates on a strict `>`).
    Returns [i, j, hh, ww] or [] (the reference's `ans`)."""
    ab = np.asarray(all_black)
    H, W = ab.shape
    run = np.zeros((H, W + 1), np.int64)
    for c in range(W - 1, -1, -1):
        run[:, c] = np.where(ab[:, c] > 0, 0, run[:, c + 1] + 1)
    max_s, ans = 0, []
    for i in range(0, H // 2, step):
        for j in range(0, W // 2, step):
            if ab[i, j] > 0:
                continue
            w = np.minimum.accumulate(run[i:, j])
            area = w * np.arange(1, H - i + 1)
            k = int(np.argmax(area))                       # first maximum
            if area[k] > max_s:
                max_s, ans = int(area[k]), [i, j, i + k, j + int(w[k]) - 1]
    return ans


# ---------------------------------------------------------------------------------------------------------------------
# warpRevBundle(img, Hs) of deploy_bundle.py:148-173 (the per-cell cv2.warpPerspective variant; its call at :300 is commented
# out in the reference, warpRevBundle2 is the live one) with cvt_theta_mat_bundle :121-134.
def _bilinear_fixed_u8(img, sx, sy):
    """remapBilinear on 1/32-pixel fixed-point coordinates (int64 arrays), constant border 0 -- shared with cv2.remap"""
    H, W = img.shape[:2]
    ix = np.clip(sx >> 5, -32768, 32767)
    iy = np.clip(sy >> 5, -32768, 32767)
    fx, fy = sx & 31, sy & 31
    w = [(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32]

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        return v * ok[..., None]

    s = (tap(iy, ix) * w[0][..., None] + tap(iy, ix + 1) * w[1][..., None] + tap(iy + 1, ix) * w[2][..., None] +
         tap(iy + 1, ix + 1) * w[3][..., None])
    return np.clip((s + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def warp_perspective_fixed_coords(M, out_w, out_h):
    """cv2.warpPerspective(..., flags=WARP_INVERSE_MAP|INTER_LINEAR), imgwarp.cpp WarpPerspectiveInvoker: the destination is
    walked in blocks (bh0 = min(16, h), bw0 = min(1024 / bh0, w), bh0 = min(1024 / bw0, h)); in a block starting at column bx,
        X0 = M0*bx + M1*y + M2, Y0 = M3*bx + M4*y + M5, W0 = M6*bx + M7*y + M8          (double, left to right)
        W = W0 + M6*x1; W = W ? 32/W : 0; fX = clamp((X0 + M0*x1)*W, INT_MIN, INT_MAX), X = cvRound(fX)   (x1 = x - bx)
    -> 1/32-pixel fixed-point source coordinates (int64 [out_h,out_w] each)."""
    M = np.asarray(M, np.float64).reshape(9)
    bh0 = min(16, out_h)
    bw0 = min(1024 // bh0, out_w)
    x = np.arange(out_w)
    bx = (x // bw0) * bw0
    x1 = (x - bx).astype(np.float64)
    bxf = bx.astype(np.float64)
    y = np.arange(out_h, dtype=np.float64)[:, None]
    X0 = M[0] * bxf[None, :] + M[1] * y + M[2]
    Y0 = M[3] * bxf[None, :] + M[4] * y + M[5]
    W0 = M[6] * bxf[None, :] + M[7] * y + M[8]
    Wv = W0 + M[6] * x1[None, :]
    with np.errstate(divide='ignore', invalid='ignore'):
        Wv = np.where(Wv != 0, 32.0 / Wv, 0.0)
        fX = np.maximum(-2147483648.0, np.minimum(2147483647.0, (X0 + M[0] * x1[None, :]) * Wv))
        fY = np.maximum(-2147483648.0, np.minimum(2147483647.0, (Y0 + M[3] * x1[None, :]) * Wv))
    return np.rint(fX).astype(np.int64), np.rint(fY).astype(np.int64)


def cvt_theta_mat_bundle(Hs, height, width, grid_h, grid_w):
    """deploy_bundle.py:121-134, the same numpy expressions"""
    from numpy.linalg import inv
    scale_mat = np.eye(3)
    scale_mat[0, 0] = width / 2.
    scale_mat[0, 2] = width / 2.
    scale_mat[1, 1] = height / 2.
    scale_mat[1, 2] = height / 2.
    Hs = np.asarray(Hs).reshape((grid_h, grid_w, 3, 3))
    return np.matmul(np.matmul(scale_mat, Hs), inv(scale_mat))


def warp_rev_bundle(img, Hs, grid_h, grid_w):
    """deploy_bundle.py:148-173: every cell of the output is cut from the frame warped by that cell's homography"""
    height, width = img.shape[:2]
    Hc = cvt_theta_mat_bundle(Hs, height, width, grid_h, grid_w)
    gh, gw = height // grid_h, width // grid_w
    out = np.zeros_like(img)
    for i in range(grid_h):
        for j in range(grid_w):
            sx, sy = warp_perspective_fixed_coords(Hc[i, j], width, height)
            full = _bilinear_fixed_u8(img, sx, sy)
            r0, r1 = i * gh, (height if i == grid_h - 1 else (i + 1) * gh)
            c0, c1 = j * gw, (width if j == grid_w - 1 else (j + 1) * gw)
            out[r0:r1, c0:c1] = full[r0:r1, c0:c1]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# cvt_img2train(img, crop_rate) of config.py:6-21: BGR uint8 frame -> cv2.cvtColor(BGR2GRAY) -> PIL resize(BILINEAR) [-> centre
# crop] -> v * (1/255) - 0.5.  Third-party arithmetic (OpenCV 4.13, Pillow 12.2 here; neither pinned by the reference), restated
# from the published algorithms and pinned against the libraries themselves (tests/golden/deploy_cvt_img2train.npz):
#   BGR2GRAY (color_rgb.simd.hpp, RGB2Gray<uchar>): (B*3735 + G*19235 + R*9798 + 2^14) >> 15
#   Pillow ImagingResample (src/libImaging/Resample.c), 8 bits per channel: per output sample a window of
#   ceil(support)*2+1 taps, support = max(scale, 1) for BILINEAR, triangle weights normalised to 1 in double, converted to
#   22-bit fixed point (round half away from zero); horizontal pass first, each pass = clip8((2^21 + sum v*k) >> 22).
def bgr2gray_u8(img):
    b, g, r = (img[..., k].astype(np.int64) for k in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def pil_bilinear_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc of Resample.c -> (xmin [out], count [out], k int32 [out, ksize])"""
    import math
    scale = in_size / float(out_size)
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, np.int32); cnt_a = np.zeros(out_size, np.int32); kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = []
        ww = 0.0
        for x in range(xmax):
            t = (x + xmin - center + 0.5) * ss
            t = -t if t < 0.0 else t
            v = 1.0 - t if t < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(v * (1 << 22) - 0.5) if v < 0 else int(v * (1 << 22) + 0.5)
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    return xmin_a, cnt_a, kk


def _pil_pass(src, xmin, cnt, kk, axis):
    """one resampling pass along `axis` of a uint8 2-D array"""
    s = np.moveaxis(src.astype(np.int64), axis, 1)           # [other, in]
    out = np.zeros((s.shape[0], len(xmin)), np.int64)
    for xx in range(len(xmin)):
        acc = np.full(s.shape[0], 1 << 21, np.int64)
        for x in range(cnt[xx]):
            acc += s[:, xmin[xx] + x] * int(kk[xx, x])
        out[:, xx] = np.clip(acc >> 22, 0, 255)
    return np.moveaxis(out.astype(np.uint8), 1, axis)


def pil_resize_bilinear_u8(gray, out_w, out_h):
    h, w = gray.shape
    t = gray
    if out_w != w:
        t = _pil_pass(t, *pil_bilinear_coeffs(w, out_w), axis=1)
    if out_h != h:
        t = _pil_pass(t, *pil_bilinear_coeffs(h, out_h), axis=0)
    return t


def cvt_img2train(img, height, width, crop_rate=1):
    """config.py:6-21 -> float64 [1,height,width,1] as the reference returns it"""
    g = bgr2gray_u8(img)
    if crop_rate != 1:
        h = int(height / crop_rate); dh = int((h - height) / 2)
        w = int(width / crop_rate); dw = int((w - width) / 2)
        g = pil_resize_bilinear_u8(g, w, h)[dh:dh + height, dw:dw + width]
    else:
        g = pil_resize_bilinear_u8(g, width, height)
    return (g * (1. / 255) - 0.5).reshape((1, height, width, 1))


# ---------------------------------------------------------------------------------------------------------------------
# cv2.resize(frame, (width, height)) of the uint8 colour frame (deploy_bundle.py:301, default INTER_LINEAR), resize.cpp:
#   coefficients as for float (fx = (dx+0.5)*scale - 0.5, border taps folded with weight 0) but in 11-bit fixed point
#   (cvRound(w * 2048)); horizontal pass in int32: S[sx]*a0 + S[sx+1]*a1; vertical pass of the 8-bit specialisation:
#   ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  An exact 2x2 decimation is rerouted by OpenCV to INTER_AREA
#   (mean of the 2x2 block, (sum + 2) >> 2).  Plain C++ path (cv2.setUseOptimized(False)); the dispatched path (IPP) may differ.
def resize_linear_u8(img, out_w, out_h):
    H, W = img.shape[:2]
    src = img.astype(np.int64).reshape(H, W, -1)
    if W == 2 * out_w and H == 2 * out_h:
        s = src[0::2, 0::2] + src[0::2, 1::2] + src[1::2, 0::2] + src[1::2, 1::2]
        return ((s + 2) >> 2).astype(np.uint8).reshape((out_h, out_w) + img.shape[2:])

    def coeffs(n_out, n_in, clamp_weights):
        scale = _scale(n_out, n_in)
        i0 = np.zeros(n_out, np.int64); i1 = np.zeros(n_out, np.int64); a = np.zeros((n_out, 2), np.int64)
        for d in range(n_out):
            f = f32((d + 0.5) * scale - 0.5)
            s0 = int(np.floor(f))
            f = f32(f - f32(s0))
            if clamp_weights:                       # horizontal: fold the outside tap onto the border with weight 0
                if s0 < 0:
                    s0, f = 0, f32(0)
                if s0 >= n_in - 1:
                    s0, f = n_in - 1, f32(0)
                i0[d], i1[d] = s0, min(s0 + 1, n_in - 1)
            else:                                   # vertical: clamp the row indices, keep the weights
                i0[d], i1[d] = min(max(s0, 0), n_in - 1), min(max(s0 + 1, 0), n_in - 1)
            a[d] = (int(np.rint(f32(f32(1) - f) * f32(2048))), int(np.rint(f * f32(2048))))
        return i0, i1, a

    x0, x1, ax = coeffs(out_w, W, True)
    y0, y1, ay = coeffs(out_h, H, False)
    hp = src[:, x0] * ax[:, 0][None, :, None] + src[:, x1] * ax[:, 1][None, :, None]          # [H, out_w, C] int
    r0, r1 = hp[y0], hp[y1]
    b0, b1 = ay[:, 0][:, None, None], ay[:, 1][:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((out_h, out_w) + img.shape[2:])


# ---------------------------------------------------------------------------------------------------------------------
# frame transport: the two literal expressions of the reference (numpy evaluates them exactly as the reference does)
def u8_to_train(frame_u8):
    """config.py:19: img * (1. / 255) - 0.5 (float64), then the float32 cast of the network's placeholder."""
    return (np.asarray(frame_u8, np.uint8) * (1. / 255) - 0.5).astype(np.float32)


def train_to_u8(x):
    """deploy_bundle.py:75 cvt_train2img without the reshape: ((x + 0.5) * 255).astype(np.uint8) on a float32 array."""
    return ((np.asarray(x, np.float32) + 0.5) * 255).astype(np.uint8)
