import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import dovs_b200 as mgw
from test_gpu_train_pass import inputs, dev, smooth_flow, composed, fused
from conftest import relmax
n, h, w, c, m = 4, 96, 128, 1, 300
raw = inputs(n, h, w, c, m, 700 + n + h)
t = {k: dev(v) for k, v in raw.items()}
flow = dev(smooth_flow(n, h, w))
Z = dict(theta_mul=0.0, grid_theta_mul=0.0, black_mul=0.0, distortion_mul=0.0, consistency_mul=0.0, img_mul=0.0, feature_mul=0.0)
for name, mul, temp in (('img only', dict(Z, img_mul=50.0), False), ('feat only', dict(Z, feature_mul=1.0), False),
                        ('vertex only', dict(Z, theta_mul=0.16, distortion_mul=1.0, consistency_mul=20.0, black_mul=120.0), False),
                        ('temp only', Z, True), ('img+temp', dict(Z, img_mul=50.0), True), ('all', None, True)):
    gs = {}
    for impl in ('auto', 'generic'):
        mgw.set_impl(impl)
        for nm, fn in (('composed', composed), ('fused', fused)):
            head = t['head'].clone().requires_grad_(True)
            total = fn(mgw, t, head, dict(mul=mul) if mul else {}, 2 * n, temp, flow)[0]
            (g,) = torch.autograd.grad(total, head)
            gs[impl, nm] = g.cpu().numpy()
    ref = gs['generic', 'composed']
    print('%-12s |g|max %.3e  ' % (name, np.abs(ref).max()), '  '.join('%s/%s %.2e' % (k[0], k[1], relmax(v, ref)) for k, v in gs.items()))
