"""Randomised agreement soak of the deploy-side kernels (GPU) against the restatements in oracle/deploy_ref.py: cvt_img2train,
cv2_resize, warpRevBundle, warpRevBundle2, crop_rect on random sizes / grids / maps.  usage: soak_deploy.py [cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
import deploy_ref as D, dovs_b200 as mgw
from dovs_b200 import ops

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
r = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
for k in range(cases):
    H, W = int(r.randint(8, 400)), int(r.randint(8, 500))
    h, w = int(r.randint(4, 300)), int(r.randint(4, 400))
    img = r.randint(0, 256, (H, W, 3)).astype(np.uint8)
    tag = (k, H, W, h, w)
    cr = float(r.choice([1.0, 0.9, 0.8]))
    got = mgw.deploy.cvt_img2train(img, 1 if cr == 1 else cr, height=h, width=w).cpu().numpy()
    assert np.array_equal(got, D.cvt_img2train(img, h, w, 1 if cr == 1 else cr).astype(np.float32)), ('cvt', tag, cr)
    c = int(r.choice([1, 3, 4]))
    imc = r.randint(0, 256, (H, W, c)).astype(np.uint8)
    if r.rand() < 0.15:
        hh, ww = H // 2, W // 2
        imc = imc[:2 * hh, :2 * ww]
        assert np.array_equal(mgw.deploy.cv2_resize(imc, (ww, hh)), D.resize_linear_u8(imc, ww, hh).reshape(hh, ww, c)), ('area2', tag)
    assert np.array_equal(mgw.deploy.cv2_resize(imc, (w, h)), D.resize_linear_u8(imc, w, h).reshape(h, w, c)), ('resize', tag, c)
    gh, gw = int(r.randint(1, min(6, H) + 1)), int(r.randint(1, min(6, W) + 1))
    Hs = (np.tile(np.eye(3, dtype=np.float32).reshape(1, 1, 9), (gh, gw, 1)) +
          float(r.choice([0.0, 0.05, 0.3])) * r.standard_normal((gh, gw, 9)).astype(np.float32) * np.array([1, 1, 1, 1, 1, 1, 0.5, 0.5, 0], np.float32)).astype(np.float32)
    assert np.array_equal(mgw.warpRevBundle(img, Hs, grid=(gh, gw)), D.warp_rev_bundle(img, Hs, gh, gw)), ('warpRevBundle', tag, gh, gw)
    if H >= 4 and W >= 4:
        x_map = (np.linspace(-1, 1, W, dtype=np.float32)[None, :] + 0.1 * r.standard_normal((H, W)).astype(np.float32)).astype(np.float32)
        y_map = (np.linspace(-1, 1, H, dtype=np.float32)[:, None] + 0.1 * r.standard_normal((H, W)).astype(np.float32)).astype(np.float32)
        assert np.array_equal(mgw.warpRevBundle2(img, x_map, y_map), D.warp_rev_bundle2(img, x_map, y_map)), ('warpRevBundle2', tag)
    ab = (r.random_sample((H, W)) < float(r.choice([0.0, 0.001, 0.02]))).astype(np.int64) * 2
    step = int(r.choice([1, 3, 10])) if H * W < 20000 else 10
    want = D.crop_rect(ab, step)
    assert ops.crop_rect(torch.as_tensor(ab.astype(np.int32)).cuda(), step).cpu().tolist() == (want if want else [-1] * 4), ('crop', tag, step)
    if k % 10 == 0:
        print(k, tag, flush=True)
print('deploy soak ok: %d cases' % cases)
