import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
import synth, dovs_b200 as mgw
n, h, w, c = 1, 1080, 1920, 3
U = synth.smooth_image(n, h, w, c, 412)
theta = synth.random_mesh(n, 4, 4, 0.03, 413)
g, gi = synth.randn((n, h, w, c), 414), synth.randn((n, h, w, 2), 415, 0.1)
d = lambda a: torch.tensor(a, device='cuda')
Hs = mgw.ops.solve_h_fwd(d(theta))
res = {}
for impl in ('generic', 'auto'):
    mgw.set_impl(impl)
    dU, dHs = mgw.ops.warp_bwd(d(U), Hs, d(g), d(gi))
    res[impl] = (dU.cpu().numpy(), dHs.cpu().numpy())
a, b = res['generic'][0], res['auto'][0]
diff = np.abs(a - b).max(axis=(0, 3))
print('max|generic|', np.abs(a).max(), 'max diff', diff.max(), 'dHs diff', np.abs(res['generic'][1] - res['auto'][1]).max() / np.abs(res['generic'][1]).max())
ys, xs = np.nonzero(diff > 1e-3 * np.abs(a).max())
print('bad px', len(ys))
if len(ys):
    print('rows', ys.min(), ys.max(), 'cols', xs.min(), xs.max())
    # histogram by 24x64 tile
    import collections
    cnt = collections.Counter((int(y) // 30, int(x) // 60) for y, x in zip(ys, xs))
    print(sorted(cnt.items())[:60])
    for y, x in list(zip(ys, xs))[:10]:
        print(y, x, a[0, y, x], b[0, y, x])
