"""Minimal driver for ncu: a few forward+backward passes of config #2 through the C ABI (no timing, no CPU leg)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
import synth, dovs_b200 as mgw
from dovs_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W, C = 288, 512, 3
dev = 'cuda'
# PROF_NOFLUSH=1: the step as bench.py runs it -- three rotating input sets (151 MB each > the 126 MB L2), no flush between
# forward and backward, dU zero-filled by the backward call itself (for `ncu --cache-control none`: DRAM bytes of the step in place)
noflush = bool(os.environ.get('PROF_NOFLUSH'))
sets = [dict(U=torch.tensor(synth.noise_image(n, H, W, C, 900 + 10 * k), device=dev), th=torch.tensor(synth.random_mesh(n, 4, 4, 0.05, 901 + 10 * k), device=dev),
             g=torch.tensor(synth.randn((n, H, W, C), 902 + 10 * k), device=dev), gi=torch.tensor(synth.randn((n, H, W, 2), 903 + 10 * k, 0.1), device=dev))
        for k in range(3 if noflush else 1)]
flush = torch.empty(64 * 1024 * 1024, device=dev)
dU_buf = torch.empty_like(sets[0]['U'])
for i in range(reps):
    s = sets[i % len(sets)]
    U, th, g, gi = s['U'], s['th'], s['g'], s['gi']
    if not noflush:
        flush.zero_()
    out, black, img, Hs = ops.mesh_warp_fwd(U, th)
    if not noflush:
        flush.zero_()
    dU, dth = ops.mesh_warp_bwd(U, th, Hs, g, gi, want_dU=not os.environ.get('PROF_NODU'), dU_out=None if os.environ.get('PROF_NODU') else dU_buf)
torch.cuda.synchronize()
print('ok', float(dth.abs().max()), mgw.launch_count())
