"""How big must the staged source box be?  For the benchmark meshes (config #2, sigma=0.05) and a tile shape, the
distribution of the (clipped) tap bounding box of every tile: which (SBW, SBH) make what fraction of tiles COMPLETE.
CPU only (numpy + the C oracle's H solve)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import synth, c_oracle

def stats(TW, TH, n=32, H=288, W=512, gh=4, gw=4, sigma=0.05, seed=901, align=4):
    th = synth.random_mesh(n, gh, gw, sigma, seed)
    Hs = c_oracle.solve_h(th, f64=True)
    ch, cw = H // gh, W // gw
    lx = np.linspace(-1, 1, W); ly = np.linspace(-1, 1, H)
    needw, needh = [], []
    for b in range(n):
        for ci in range(gh):
            for cj in range(gw):
                Hm = Hs[b, ci, cj].reshape(3, 3)
                for r0 in range(ci * ch, (ci + 1) * ch, TH):
                    r0 = min(r0, (ci + 1) * ch - TH)
                    for c0 in range(cj * cw, (cj + 1) * cw, TW):
                        c0 = min(c0, (cj + 1) * cw - TW)
                        xs, ys = [], []
                        for rr in (r0, r0 + TH - 1):
                            for cc in (c0, c0 + TW - 1):
                                v = Hm @ np.array([lx[cc], ly[rr], 1.0])
                                xs.append((v[0] / v[2] + 1) * W / 2); ys.append((v[1] / v[2] + 1) * H / 2)
                        ux0, ux1 = int(np.floor(min(xs))) - 1, int(np.floor(max(xs))) + 2
                        uy0, uy1 = int(np.floor(min(ys))) - 1, int(np.floor(max(ys))) + 2
                        ix0, ix1 = np.clip(ux0, 0, W - 1), np.clip(ux1, 0, W - 1)
                        iy0, iy1 = np.clip(uy0, 0, H - 1), np.clip(uy1, 0, H - 1)
                        bx0 = ix0 - ix0 % align
                        needw.append(ix1 - bx0 + 1); needh.append(iy1 - iy0 + 1)
    return np.array(needw), np.array(needh)

if __name__ == '__main__':
    for TW, TH in [(64, 24), (64, 12), (32, 24), (32, 16), (32, 32), (64, 18), (32, 12), (32, 18), (32,36)]:
        w, h = stats(TW, TH)
        q = lambda a: ' '.join('%d' % np.percentile(a, p) for p in (50, 90, 95, 99, 99.9, 100))
        print('tile %2dx%2d (WxH): need w [p50 p90 p95 p99 p99.9 max] = %s | need h = %s' % (TW, TH, q(w), q(h)))
