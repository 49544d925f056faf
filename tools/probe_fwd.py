"""Phase timing of the forward pipeline's consumers (debug build with -DMGW_PROBE): mean SM cycles per tile between probe points."""
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
flush = torch.empty(64 * 1024 * 1024, device=dev)
buf = (ctypes.c_ulonglong * 16)()
names = ['wait for the records', 'phase 1: record read + projective map', 'phase 2: map / mask stores', 'phase 3: taps', 'wait for the box',
         'phase 4: gather', 'release + proxy fence + wait for the previous store', 'consumer barrier', 'store issue']
for rep in range(3):
    flush.zero_()
    lib.mgw_debug_probe_fwd(buf, 1)
    ops.mesh_warp_fwd(U, th)
    lib.mgw_debug_probe_fwd(buf, 0)
    tiles = buf[15]
    tot = sum(buf[i] for i in range(9))
    print('rep %d: %d tiles seen by the probed thread, mean %.0f cycles per tile' % (rep, tiles, tot / max(tiles, 1)))
    for i, nm in enumerate(names):
        print('  %-58s %7.0f cycles  %5.1f %%' % (nm, buf[i] / max(tiles, 1), 100.0 * buf[i] / max(tot, 1)))
