import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs /root/reference (build container only); skipped elsewhere')


def pytest_collection_modifyitems(config, items):
    import ref_loader
    if not ref_loader.available():
        skip = pytest.mark.skip(reason='/root/reference not present on this box')
        for it in items:
            if 'reference' in it.keywords:
                it.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def golden_inputs(name, g):
    """(U, d_out, d_img) of a mesh fixture; the full-size ones regenerate them from the synth seeds."""
    import synth
    if 'U' in g:
        return g['U'], g['d_out'], g['d_img']
    n, h, w, c = (int(v) for v in g['shape'])
    seed = int(g['seed'])
    U = synth.noise_image(n, h, w, c, seed) if str(g['kind']) == 'noise' else synth.smooth_image(n, h, w, c, seed)
    return U, synth.randn((n, h, w, c), seed + 100), synth.randn((n, h, w, 2), seed + 200, 0.1)


def golden_black(g):
    n, h, w, _ = (int(v) for v in g['shape'])
    return np.unpackbits(g['ref_black_bits'])[:n * h * w].reshape(n, h, w).astype(np.float32)


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return (a.view(np.int32) == b.view(np.int32)) | (np.isnan(a) & np.isnan(b))


SMALL_MESH_CASES = ['mesh_smooth_s03', 'mesh_noise_s08', 'mesh_identity', 'mesh_ragged_c1', 'mesh_grid23_c4',
                    'mesh_fold_clamp_shift']
TMA_MESH_CASES = ['mesh_tma_ragged', 'mesh_tma_fold_clamp_shift', 'mesh_tma_c1', 'mesh_tma_c4_g22']     # cells >= 8 x 32 px
FULL_MESH_CASES = ['mesh_full_noise_s05', 'mesh_full_identity']
MESH_CASES = SMALL_MESH_CASES + TMA_MESH_CASES + FULL_MESH_CASES


def relmax(a, b):
    """max|a-b| / max|b|: the 'relative' of BASELINE.json's gradient tolerance (hard part 2)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
