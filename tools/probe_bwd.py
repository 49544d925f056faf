"""Phase timing of the tile backward (debug build with -DMGW_PROBE): mean SM cycles per CTA between probe points."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, synth, dovs_b200 as mgw
from dovs_b200 import ops
from dovs_b200._lib import lib
n, H, W, C = 32, 288, 512, 3
dev = 'cuda'
U = torch.tensor(synth.noise_image(n, H, W, C, 900), device=dev)
th = torch.tensor(synth.random_mesh(n, 4, 4, 0.05, 901), device=dev)
g = torch.tensor(synth.randn((n, H, W, C), 902), device=dev)
gi = torch.tensor(synth.randn((n, H, W, 2), 903, 0.1), device=dev)
out, black, img, Hs = ops.mesh_warp_fwd(U, th)
dU_buf = torch.empty_like(U)
flush = torch.empty(64 * 1024 * 1024, device=dev)
buf = (ctypes.c_ulonglong * 16)()
names = ['gradients staged by TMA: CTA barrier + wait + copy to registers', 'max|d_out|, barrier, zero the accumulator', 'barrier 1', 'scale',
         'wait for the TMA box of U', 'pixel loop', 'dH shuffles', 'barrier 2', 'drain: fixed point -> fp32 -> red.global.add.v4.f32', '(unused)', '(unused)']
for rep in range(3):
    flush.zero_()
    lib.mgw_debug_probe(buf, 1)
    ops.mesh_warp_bwd(U, th, Hs, g, gi, want_dU=True, dU_out=dU_buf)
    lib.mgw_debug_probe(buf, 0)
    ctas = buf[15]
    tot = sum(buf[i] for i in range(14))
    print('rep %d: %d CTAs, mean lifetime %.0f cycles' % (rep, ctas, tot / max(ctas, 1)))
    for i, nm in [(11, 'barrier init, gradient TMA requests (thread 0)'), (12, 'tile decode'), (13, 'Hs loads; warp 0: source box + TMA box request')] + list(enumerate(names)):
        print('  %-58s %7.0f cycles  %5.1f %%' % (nm, buf[i] / max(ctas, 1), 100.0 * buf[i] / max(tot, 1)))
