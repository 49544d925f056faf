"""TEST INFRASTRUCTURE -- ctypes front end of oracle/mgw_oracle.c (the bit-exact CPU checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'mgw_oracle.c')
SO = os.path.join(HERE, '_build', 'libmgw_oracle.so')

_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off: no implicit FMA contraction, the rounding order is the source's."""
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-std=c11', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared',
                               '-o', SO, SRC, '-lm'])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def vertices(head, gh, gw, do_crop_rate=0.8):
    head = _f(head)
    n = head.shape[0]
    pts2 = np.empty((n, gh + 1, gw + 1, 2), np.float32)
    pts1 = np.empty((n, gh, gw, 8), np.float32)
    lib().orc_vertices(_p(head), n, gh, gw, ctypes.c_float(do_crop_rate), _p(pts2), _p(pts1))
    return pts1, pts2


def solve_h(theta, f64=False):
    theta = _f(theta)
    n, gh1, gw1, _ = theta.shape
    Hs = np.empty((n, gh1 - 1, gw1 - 1, 9), np.float64 if f64 else np.float32)
    (lib().orc_solve_h_f64 if f64 else lib().orc_solve_h)(_p(theta), n, gh1 - 1, gw1 - 1, _p(Hs))
    return Hs


def warp(U, Hs, want_out=True):
    U, Hs = _f(U), _f(Hs)
    n, h, w, c = U.shape
    gh, gw = Hs.shape[1:3]
    out = np.empty((n, h, w, c), np.float32) if want_out else None
    black = np.empty((n, h, w), np.float32)
    img = np.empty((n, h, w, 2), np.float32)
    cell = np.empty((n, h, w), np.int32)
    lib().orc_warp(_p(U), _p(Hs), n, h, w, c, gh, gw, _p(out), _p(black), _p(img), _p(cell))
    return out, black, img, cell


def interp(im, x, y, out_size):
    im, x, y = _f(im), _f(x), _f(y)
    n, ih, iw, c = im.shape
    oh, ow = out_size
    assert x.size == n * oh * ow and y.size == n * oh * ow
    out = np.empty((n, oh, ow, c), np.float32)
    lib().orc_interp(_p(im), _p(x), _p(y), n, ih, iw, c, oh, ow, _p(out))
    return out


def homography_warp(U, theta, out_size):
    U, theta = _f(U), _f(theta).reshape(-1, 9)
    n, h, w, c = U.shape
    oh, ow = out_size
    out = np.empty((n, oh, ow, c), np.float32)
    black = np.empty((n, oh, ow), np.float32)
    img = np.empty((n, oh, ow, 2), np.float32)
    lib().orc_homography_warp(_p(U), _p(theta), n, h, w, c, oh, ow, _p(out), _p(black), _p(img))
    return out, black, img
