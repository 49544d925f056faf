"""CUDA-event timing of every remaining entry point of the path at config #2 size (N=32, 288x512; C=3 frames, C=1 for the
reference's own gray frames): interpolate, the loss epilogues, the fused warp+img_loss pair, the vertex builder.
Bytes are the algorithmic ones (each tensor read / written once).  Not part of the bench.py contract."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
import synth, dovs_b200 as mgw
from dovs_b200 import ops
dev = 'cuda'
n, H, W = 32, 288, 512
flush = torch.empty(48 * 1024 * 1024, device=dev)
res = {}


def timeit(name, fn, nbytes, reps=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    res[name] = {'us': round(us, 1), 'algorithmic_MB': round(nbytes / 1e6, 1), 'GB_per_s': round(nbytes / us / 1e3, 0)}
    print('%-44s %8.1f us  %7.1f MB  %6.0f GB/s' % (name, us, nbytes / 1e6, nbytes / us / 1e3))


for C in (3, 1):
    P = n * H * W
    U = torch.tensor(synth.noise_image(n, H, W, C, 1), device=dev)
    y = torch.tensor(synth.noise_image(n, H, W, C, 2), device=dev)
    th = torch.tensor(synth.random_mesh(n, 4, 4, 0.05, 3), device=dev)
    out, black, img, Hs = ops.mesh_warp_fwd(U, th)
    g = torch.tensor(synth.randn((n, H, W, C), 4), device=dev)
    x_flow = (img[..., 0:1] + 0.01).contiguous(); y_flow = (img[..., 1:2] - 0.01).contiguous()
    tag = 'C%d ' % C
    timeit(tag + 'interp_fwd', lambda: ops.interp_fwd(U, x_flow, y_flow, (H, W)), P * (8 * C + 8))
    timeit(tag + 'interp_bwd (d_im + dx,dy)', lambda: ops.interp_bwd(U, x_flow, y_flow, g, (H, W)), P * (12 * C + 16))
    timeit(tag + 'img_loss_fwd', lambda: ops.img_loss_fwd(out, y, black), P * (8 * C + 4))
    sums = ops.img_loss_fwd(out, y, black)
    timeit(tag + 'img_loss_bwd', lambda: ops.img_loss_bwd(out, y, black, sums, 1.0), P * (12 * C + 4))
    timeit(tag + 'mesh_warp_img_loss_fwd (fused)', lambda: ops.mesh_warp_img_loss_fwd(U, th, y), P * (12 * C + 12))
    o2, b2, i2, Hs2, sums2 = ops.mesh_warp_img_loss_fwd(U, th, y)
    timeit(tag + 'mesh_warp_img_loss_bwd (fused, dU + dtheta)', lambda: ops.mesh_warp_img_loss_bwd(U, th, Hs2, o2, y, b2, sums2, 1.0, float(n)),
           P * (16 * C + 4))
    timeit(tag + 'temp_loss_fwd', lambda: ops.temp_loss_fwd(out, black, o2, b2, img), P * (8 * C + 16))
    ts = ops.temp_loss_fwd(out, black, o2, b2, img)
    timeit(tag + 'temp_loss_bwd', lambda: ops.temp_loss_bwd(out, black, o2, b2, img, ts, 1.0), P * (16 * C + 16))
M = 3000
matches = torch.tensor(synth.uniform((n, M, 4), -1, 1, 5), device=dev)
mask = (torch.rand(n, M, device=dev) < 0.3).float()
timeit('feature_loss_fwd (3000 matches/sample)', lambda: ops.feature_loss_fwd(matches, mask, img), n * M * 28)
timeit('feature_loss_bwd', lambda: ops.feature_loss_bwd(matches, mask, img, 1.0), n * M * 28 + P * 8)
_Hs_feat = ops.solve_h_fwd(torch.tensor(synth.random_mesh(n, 4, 4, 0.05, 77), device=dev))
timeit('feature_loss_dh (backward straight to dH partials; this wrapper also forms the mask counts with 4 torch kernels: the kernel itself is 11 us)', lambda: ops.feature_loss_dh(matches, mask, img, _Hs_feat, 1.0), n * M * 28)
head = torch.tensor(synth.randn((n, 50), 6, 0.05), device=dev)
timeit('vertices_fwd (get_4_pts)', lambda: ops.vertices_fwd(head, 4, 4), n * 50 * 12)
pts1, pts2 = ops.vertices_fwd(head, 4, 4)
f4 = torch.ones(4, device=dev)
timeit('vertex_losses_fwd (id+black_pos+distortion+consistency)', lambda: ops.vertex_losses_fwd(head, pts1, pts2, 4, 4), n * (50 + 128 + 50) * 4)
timeit('vertex_losses_bwd', lambda: ops.vertex_losses_bwd(head, pts1, pts2, 4, 4, f4), n * (50 + 128 + 50) * 8)
# deploy-side crop (deploy_bundle.py:291,344-365) at the network's frame size: accumulation per frame, search once per video
blk = (torch.rand(H, W, device=dev) < 0.03).float()
ab = torch.zeros(H, W, device=dev, dtype=torch.int32)
timeit('black_accumulate 288x512', lambda: ops.black_accumulate(ab, blk), H * W * 12)
yy, xx = np.mgrid[0:H, 0:W]
border = ((yy < 12 + 8 * np.sin(xx / 37.0)) | (xx < 18 + 10 * np.cos(yy / 23.0)) | (yy > H - 14 - 6 * np.sin(xx / 51.0)) |
          (xx > W - 16 + 8 * np.sin(yy / 29.0))).astype(np.int32)
abv = torch.tensor(border, device=dev)
timeit('crop_rect 288x512 (390 corners)', lambda: ops.crop_rect(abv), H * W * 8)
import time, deploy_ref
t0 = time.perf_counter(); want = deploy_ref.crop_rect(border.astype(np.int64)); t_np = time.perf_counter() - t0
assert ops.crop_rect(abv).cpu().tolist() == want
res['crop_rect 288x512 (390 corners)']['numpy_restatement_us'] = round(t_np * 1e6, 0)
res['crop_rect 288x512 (390 corners)']['rect'] = want
print(json.dumps(res))
