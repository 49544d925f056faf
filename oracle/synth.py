"""TEST INFRASTRUCTURE -- deterministic synthetic inputs shared by the golden generator, the tests and bench.py.

Pure numpy (no torch RNG) so the same arrays are reproduced on any box: SURVEY.md 8(d).
"""
import numpy as np


def smooth_image(n, h, w, c, seed=0):
    """band-limited image, |grad| <~ 1e-2 per pixel: 0.4*sin(x/37+y/53+ch)*cos(y/41-ch), per-sample shift."""
    rng = np.random.RandomState(seed)
    y = np.arange(h, dtype=np.float64)[None, :, None, None]
    x = np.arange(w, dtype=np.float64)[None, None, :, None]
    ch = np.arange(c, dtype=np.float64)[None, None, None, :]
    sh = rng.uniform(0, 50, size=(n, 1, 1, 1))
    img = 0.4 * np.sin((x + sh) / 37.0 + y / 53.0 + ch) * np.cos((y - sh) / 41.0 - ch)
    return img.astype(np.float32)


def noise_image(n, h, w, c, seed=0):
    """white noise in [-0.5, 0.5): the reference's img/255-0.5 range (config.py:19)."""
    rng = np.random.RandomState(seed)
    return (rng.random_sample((n, h, w, c)) - 0.5).astype(np.float32)


def identity_mesh(n, gh, gw):
    """vertex (i,j) = (-1+2j/gw, -1+2i/gh), (x,y) last: s_net_bundle_nobm.py:44-46."""
    ys = (np.arange(gh + 1, dtype=np.float64) * (2.0 / gh) - 1).astype(np.float32)
    xs = (np.arange(gw + 1, dtype=np.float64) * (2.0 / gw) - 1).astype(np.float32)
    m = np.zeros((n, gh + 1, gw + 1, 2), np.float32)
    m[..., 0] = xs[None, None, :]
    m[..., 1] = ys[None, :, None]
    return m


def random_mesh(n, gh, gw, sigma, seed=1, clamp=1.25):
    rng = np.random.RandomState(seed)
    m = identity_mesh(n, gh, gw) + (sigma * rng.standard_normal((n, gh + 1, gw + 1, 2))).astype(np.float32)
    return np.clip(m, -clamp, clamp).astype(np.float32)


def randn(shape, seed, scale=1.0):
    return (scale * np.random.RandomState(seed).standard_normal(shape)).astype(np.float32)


def uniform(shape, lo, hi, seed):
    return np.random.RandomState(seed).uniform(lo, hi, size=shape).astype(np.float32)


def random_homography(n, sigma, seed=5):
    """[N,9] near-identity homographies with H[8] != 1 so the /H[2,2] normalisation matters."""
    rng = np.random.RandomState(seed)
    h = np.tile(np.eye(3, dtype=np.float64).reshape(1, 9), (n, 1))
    h += sigma * rng.standard_normal((n, 9))
    h *= rng.uniform(0.8, 1.3, size=(n, 1))
    return h.astype(np.float32)
