"""config #5's path without the backbone (theta is a leaf): one eager pass after warm-up, for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import synth, dovs_b200 as mgw
dev = 'cuda'
nb = int(os.environ.get('NB', 32))
C = int(os.environ.get('CH', 1))


def clip_batch(seed):
    return dict(x=torch.tensor(synth.noise_image(nb, 288, 512, C, seed), device=dev),
                y=torch.tensor(synth.noise_image(nb, 288, 512, C, seed + 1), device=dev),
                matches=torch.tensor(synth.uniform((nb, 3000, 4), -1, 1, seed + 2), device=dev),
                mask=(torch.rand(nb, 3000, device=dev) < 0.3).float())


b1, b2 = clip_batch(40), clip_batch(50)
_ys, _xs = torch.meshgrid(torch.linspace(-1, 1, 288, device=dev), torch.linspace(-1, 1, 512, device=dev), indexing='ij')
flow = torch.stack([_xs + 0.03 * torch.sin(3 * _ys), _ys + 0.03 * torch.cos(2 * _xs)], -1).expand(nb, 288, 512, 2).contiguous()
theta_leaf = torch.tensor(synth.randn((nb, 50), 3, 0.03), device=dev)


def path_only():
    th = theta_leaf.detach().requires_grad_(True)
    tot = 0
    rets = []
    for b in (b1, b2):
        if os.environ.get('FUSED', '1') == '1':
            t, _, out, black, _, _, _ = mgw.train_pass(th, b['x'], b['y'], b['matches'], b['mask'], batch_size=256)
        else:
            p1, p2 = mgw.get_4_pts(th, grid=(4, 4))
            il, out, black, fl = mgw.transformer_img_loss(b['x'], p2, b['y'], batch_size=256)
            ftl, _ = mgw.feature_loss(b['matches'], b['mask'], fl, batch_size=256)
            t, _ = mgw.total_loss(th, p1, p2, il, ftl, batch_size=256)
        tot = tot + t
        rets.append((out, black))
    tot = tot + 500.0 * mgw.temp_loss(rets[0][0], rets[0][1], rets[1][0], rets[1][1], flow, batch_size=256)
    tot.backward()
    return tot


for _ in range(3):
    path_only()
torch.cuda.synchronize()
if os.environ.get('GRAPH'):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        path_only()
    for _ in range(3):
        g.replay()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print('graph replay us', ts[len(ts) // 2])
else:
    torch.cuda.profiler.start()
    path_only()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('ok')
