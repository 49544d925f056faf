"""temp_loss tile kernels vs the per-pixel kernels (MGW_LOSS_TILE=0) and vs the fp64 port; timing at config #2 size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import synth, dovs_b200 as mgw
from dovs_b200 import ops
from test_gpu_configs import _loss_inputs
dev = lambda a: torch.tensor(a, device='cuda')
def timeit(fn, reps=20):
    flush = torch.empty(48 * 1024 * 1024, device='cuda')
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))
for (n, h, w, c) in ((32, 288, 512, 3), (32, 288, 512, 1), (3, 100, 132, 4)):
    o1, o2, y, b1, b2, flow = _loss_inputs(n, h, w, c, 77 + c)
    a = [dev(x) for x in (o1, b1, o2, b2, flow)]
    sums = ops.temp_loss_fwd(*a)
    d1, d2 = ops.temp_loss_bwd(*a, sums, 1.0)
    print((n, h, w, c), 'fwd %.1f us  bwd %.1f us' % (timeit(lambda: ops.temp_loss_fwd(*a)), timeit(lambda: ops.temp_loss_bwd(*a, sums, 1.0))), flush=True)
