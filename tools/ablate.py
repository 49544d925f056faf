"""CUDA-event timing of the warp entry points with outputs switched off one at a time (where does the time go?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
import synth, dovs_b200 as mgw
from dovs_b200 import ops
from dovs_b200._lib import lib, check
n, H, W, C = 32, 288, 512, 3
dev = 'cuda'
U = torch.tensor(synth.noise_image(n, H, W, C, 900), device=dev)
th = torch.tensor(synth.random_mesh(n, 4, 4, float(os.environ.get('SIGMA', '0.05')), 901), device=dev)
g = torch.tensor(synth.randn((n, H, W, C), 902), device=dev)
gi = torch.tensor(synth.randn((n, H, W, 2), 903, 0.1), device=dev)
Hs = ops.solve_h_fwd(th)
flush = torch.empty(48 * 1024 * 1024, device=dev)
out = torch.empty_like(U); black = torch.empty((n, H, W), device=dev); img = torch.empty((n, H, W, 2), device=dev)
dU = torch.empty_like(U); dHs = torch.empty_like(Hs); dth = torch.empty_like(th)
wsb = lib.mgw_mesh_warp_bwd_workspace_bytes(n, H, W, C, 4, 4)
ws = torch.empty(wsb // 4 + 64, device=dev)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: None if t is None else t.data_ptr()

def timeit(name, fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print('%-52s %7.1f us (min %7.1f)' % (name, float(np.median(ts)), min(ts)))

fw = lambda o, b, i: check(lib.mgw_warp_fwd(P(U), P(Hs), n, H, W, C, 4, 4, P(o), P(b), P(i), None, st), 'fwd')
bw = lambda du, di: check(lib.mgw_warp_bwd(P(U), P(Hs), P(g), P(di), n, H, W, C, 4, 4, P(du), P(dHs), P(ws), st), 'bwd')
timeit('fwd: out + black + img', lambda: fw(out, black, img))
timeit('fwd: out only', lambda: fw(out, None, None))
timeit('fwd: black + img only (no gather, no TMA)', lambda: fw(None, black, img))
timeit('bwd: dU + dHs, with d_img', lambda: bw(dU, gi))
timeit('bwd: dU + dHs, no d_img', lambda: bw(dU, None))
timeit('bwd: dHs only (no scatter), with d_img', lambda: bw(None, gi))
timeit('memset dU (torch zero_)', lambda: dU.zero_())
timeit('solve_h_fwd', lambda: ops.solve_h_fwd(th))
timeit('solve_h_bwd', lambda: ops.solve_h_bwd(th, Hs, dHs))
timeit('mesh fwd (K1+K2)', lambda: check(lib.mgw_mesh_warp_fwd(P(U), P(th), n, H, W, C, 4, 4, P(Hs), P(out), P(black), P(img), st), 'mf'))
timeit('mesh bwd (memsets+K3+K4)', lambda: check(lib.mgw_mesh_warp_bwd(P(U), P(th), P(Hs), P(g), P(gi), n, H, W, C, 4, 4, P(dU), P(dth), P(ws), st), 'mb'))
mgw.set_impl('tma')
timeit('tile (one TMA tile per CTA) fwd', lambda: fw(out, black, img))
timeit('tile bwd: dU + dHs, with d_img', lambda: bw(dU, gi))
if os.environ.get('ABLATE_GENERIC'):
    mgw.set_impl('generic')
    timeit('generic fwd', lambda: fw(out, black, img))
    timeit('generic bwd', lambda: bw(dU, gi))
