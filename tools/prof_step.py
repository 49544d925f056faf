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
U = torch.tensor(synth.noise_image(n, H, W, C, 900), device=dev)
th = torch.tensor(synth.random_mesh(n, 4, 4, 0.05, 901), device=dev)
g = torch.tensor(synth.randn((n, H, W, C), 902), device=dev)
gi = torch.tensor(synth.randn((n, H, W, 2), 903, 0.1), device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for _ in range(reps):
    flush.zero_()
    out, black, img, Hs = ops.mesh_warp_fwd(U, th)
    flush.zero_()
    dU, dth = ops.mesh_warp_bwd(U, th, Hs, g, gi, want_dU=not os.environ.get('PROF_NODU'))
torch.cuda.synchronize()
print('ok', float(dth.abs().max()), mgw.launch_count())
