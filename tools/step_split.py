#!/usr/bin/env python
"""Where does the fixed cost of the step go?  CUDA-graph replays of sub-sets of the step's kernels (K1 solve, K2 forward,
fill + K3 backward, K4 solve-backward) at several batch sizes.  Timing aid only (the sub-set graphs compute incomplete results).

    python tools/step_split.py [batch ...]
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module('deep-online-video-stabilization_b200')
ops = pkg.ops
lib = pkg._lib.lib
_p = ops._p


def main():
    batches = [int(a) for a in sys.argv[1:]] or [32, 64, 128]
    dev = torch.device('cuda:0')
    H, W, C, G = 288, 512, 3, 4
    for n in batches:
        nset = max(3, (3 * 32) // n)
        g = torch.Generator(device='cpu').manual_seed(1)
        sets = []
        for _ in range(nset):
            ident = torch.stack(torch.meshgrid(torch.linspace(-1, 1, G + 1), torch.linspace(-1, 1, G + 1), indexing='ij')[::-1], -1)
            theta = (ident[None] + 0.05 * torch.randn(n, G + 1, G + 1, 2, generator=g)).float()
            sets.append(dict(U=torch.rand(n, H, W, C, generator=g).to(dev), theta=theta.to(dev),
                             d_out=torch.randn(n, H, W, C, generator=g).to(dev), d_img=(0.1 * torch.randn(n, H, W, 2, generator=g)).to(dev)))
        dU = torch.empty_like(sets[0]['U'])
        dth = torch.empty_like(sets[0]['theta'])
        Hs_fix = [ops.solve_h_fwd(s['theta']) for s in sets]
        ws = ops._workspace(lib.mgw_mesh_warp_bwd_workspace_bytes(n, H, W, C, G, G), dev)
        st = ops._st

        def run(i, k1, k4, fwd=True, bwd=True):
            s = sets[i % nset]
            if fwd:
                if k1:
                    out, black, img, Hs = ops.mesh_warp_fwd(s['U'], s['theta'])
                else:
                    Hs = Hs_fix[i % nset]
                    ops.warp_fwd(s['U'], Hs)
            else:
                Hs = Hs_fix[i % nset]
            if bwd:
                if k4:
                    ops.mesh_warp_bwd(s['U'], s['theta'], Hs, s['d_out'], s['d_img'], dU_out=dU, dtheta_out=dth)
                else:
                    pkg._lib.check(lib.mgw_warp_bwd(_p(s['U']), _p(Hs), _p(s['d_out']), _p(s['d_img']), n, H, W, C, G, G, _p(dU), None,
                                                   _p(ws), st()), 'bwd')

        variants = {
            'full  K1 K2 fill K3 K4': dict(k1=True, k4=True),
            'no K1    K2 fill K3 K4': dict(k1=False, k4=True),
            'no K4 K1 K2 fill K3   ': dict(k1=True, k4=False),
            'core     K2 fill K3   ': dict(k1=False, k4=False),
            'fwd   K1 K2           ': dict(k1=True, k4=False, bwd=False),
            'fwd      K2           ': dict(k1=False, k4=False, bwd=False),
            'bwd        fill K3 K4 ': dict(k1=False, k4=True, fwd=False),
            'bwd        fill K3    ': dict(k1=False, k4=False, fwd=False),
        }
        print('batch %d' % n)
        for name, kw in variants.items():
            for per_graph in (1, 2):
                for i in range(nset):
                    run(i, **kw)
                torch.cuda.synchronize()
                graphs = []
                for i in range(0, nset * per_graph, per_graph):
                    gph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gph):
                        for j in range(per_graph):
                            run(i + j, **kw)
                    graphs.append(gph)
                torch.cuda.synchronize()
                for i in range(10):
                    graphs[i % len(graphs)].replay()
                reps = 60
                best = []
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    for i in range(reps):
                        graphs[i % len(graphs)].replay()
                    e1.record()
                    torch.cuda.synchronize()
                    best.append(e0.elapsed_time(e1) * 1e3 / (reps * per_graph))
                best.sort()
                print('  %s  steps/graph %d : %7.1f us/step  (%.1f us per 32 frames)' % (name, per_graph, best[2], best[2] * 32 / n))
        sys.stdout.flush()


if __name__ == '__main__':
    main()
