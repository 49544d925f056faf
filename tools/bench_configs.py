"""Numbers for the BASELINE.json configs that are not the bench.py line (they are parity cases there):
  #1 single 288x512x3 frame forward, #3 StabNet-v2_93-shaped forward (torch ResNet-50 as carrier, random init, batch 16)
  with the multi-grid warp stage on our kernels, #4 1080p batch-1 streaming warp latency (p50 over 1000 frames).
Writes one JSON object to stdout.  Not part of the bench.py contract."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch, torch.nn as nn
import synth, dovs_b200 as mgw
from dovs_b200 import ops
from dovs_b200._lib import lib, check
dev = 'cuda'
res = {}
flush = torch.empty(40 * 1024 * 1024, device=dev)


def dev_time(fn, reps=50, flush_l2=True):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


# ---- config #1: single frame forward
U = torch.tensor(synth.noise_image(1, 288, 512, 3, 1), device=dev)
th = torch.tensor(synth.random_mesh(1, 4, 4, 0.05, 2), device=dev)
t = dev_time(lambda: ops.mesh_warp_fwd(U, th))
res['config1_single_frame_fwd'] = {'us_device': t, 'mpix_per_s': 288 * 512 / t}

# ---- config #4: 1080p streaming, batch 1: preallocated buffers, one CUDA graph (K1 + K2), pinned host frames
Hh, Ww = 1080, 1920
frame_h = torch.tensor(synth.noise_image(1, Hh, Ww, 3, 3)).pin_memory()
th = torch.tensor(synth.random_mesh(1, 4, 4, 0.03, 4), device=dev)
frame_d = torch.empty((1, Hh, Ww, 3), device=dev)
out_d = torch.empty_like(frame_d); black_d = torch.empty((1, Hh, Ww), device=dev); img_d = torch.empty((1, Hh, Ww, 2), device=dev)
Hs_d = torch.empty((1, 4, 4, 9), device=dev)
out_h = torch.empty((1, Hh, Ww, 3)).pin_memory()
P = lambda x: x.data_ptr()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    call = lambda: check(lib.mgw_mesh_warp_fwd(P(frame_d), P(th), 1, Hh, Ww, 3, 4, 4, P(Hs_d), P(out_d), P(black_d), P(img_d),
                                                torch.cuda.current_stream().cuda_stream), 'fwd')
    frame_d.copy_(frame_h, non_blocking=True)
    for _ in range(3):
        call()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        call()
    dev_eager = dev_time(call, 200, flush_l2=False)
    dev_graph = dev_time(g.replay, 200, flush_l2=False)
    lat = []
    for i in range(1000):
        t0 = time.perf_counter()
        frame_d.copy_(frame_h, non_blocking=True)
        g.replay()
        out_h.copy_(out_d, non_blocking=True)
        s.synchronize()
        lat.append((time.perf_counter() - t0) * 1e6)
    lat2 = []
    for i in range(1000):
        t0 = time.perf_counter()
        g.replay()
        s.synchronize()
        lat2.append((time.perf_counter() - t0) * 1e6)
res['config4_1080p_stream'] = {
    'us_device_eager_call': dev_eager, 'us_device_graph_replay': dev_graph,
    'us_p50_host_to_host_fp32_frames': float(np.percentile(lat, 50)), 'us_p99_host_to_host': float(np.percentile(lat, 99)),
    'us_p50_launch_to_done_resident': float(np.percentile(lat2, 50)),
    'bytes_h2d_per_frame': frame_h.numel() * 4, 'bytes_d2h_per_frame': out_h.numel() * 4,
    'note': 'host-to-host is PCIe-bound: 24.9 MB each way per fp32 1080p frame'}


# ---- deploy-style per-frame loop at network size (deploy_bundle.py:285-303): the 1-channel current frame goes through the
# operator (K1 + K2, C = 1), its x/y maps warp the uint8 colour frame (mgw_remap_bundle_u8); uint8 frames travel over PCIe
hn, wn = 288, 512
gray_h = torch.tensor(synth.noise_image(1, hn, wn, 1, 7)).pin_memory()
col_h = torch.randint(0, 256, (1, hn, wn, 3), dtype=torch.uint8).pin_memory()
gray_d = torch.empty((1, hn, wn, 1), device=dev); col_d = torch.empty((1, hn, wn, 3), device=dev, dtype=torch.uint8)
o1 = torch.empty_like(gray_d); b1 = torch.empty((1, hn, wn), device=dev); xy1 = torch.empty((1, hn, wn, 2), device=dev)
Hs1 = torch.empty((1, 4, 4, 9), device=dev); dst_d = torch.empty_like(col_d); dst_h = torch.empty((1, hn, wn, 3), dtype=torch.uint8).pin_memory()
ws1 = torch.empty(lib.mgw_remap_bundle_u8_workspace_bytes(1, hn, wn) // 4 + 2, device=dev)
with torch.cuda.stream(s):
    def frame_call():
        st = torch.cuda.current_stream().cuda_stream
        check(lib.mgw_mesh_warp_fwd(P(gray_d), P(th), 1, hn, wn, 1, 4, 4, P(Hs1), P(o1), P(b1), P(xy1), st), 'fwd')
        check(lib.mgw_remap_bundle_u8(P(col_d), P(xy1), 1, hn, wn, 3, P(dst_d), P(ws1), st), 'remap')
    gray_d.copy_(gray_h, non_blocking=True); col_d.copy_(col_h, non_blocking=True)
    for _ in range(3):
        frame_call()
    s.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2, stream=s):
        frame_call()
    dev_frame = dev_time(g2.replay, 200, flush_l2=False)
    dev_remap = dev_time(lambda: check(lib.mgw_remap_bundle_u8(P(col_d), P(xy1), 1, hn, wn, 3, P(dst_d), P(ws1),
                                                               torch.cuda.current_stream().cuda_stream), 'remap'), 200, flush_l2=False)
    lat3 = []
    for i in range(1000):
        t0 = time.perf_counter()
        gray_d.copy_(gray_h, non_blocking=True); col_d.copy_(col_h, non_blocking=True)
        g2.replay()
        dst_h.copy_(dst_d, non_blocking=True)
        s.synchronize()
        lat3.append((time.perf_counter() - t0) * 1e6)
# the other deploy warp, warpRevBundle (deploy_bundle.py:148-173): 16 cv2.warpPerspective calls + stitching on the CPU
dev_wrb = cpu_wrb = None
try:
    import cv2
    Hs_np = Hs1.cpu().numpy()[0]                      # the operator's Hs of the frame above
    Hc_np = mgw.deploy.cvt_theta_mat_bundle(Hs_np, hn, wn, 4, 4)
    Hc_d = torch.as_tensor(np.ascontiguousarray(Hc_np)).cuda().reshape(1, 4, 4, 9)
    with torch.cuda.stream(s):
        dev_wrb = dev_time(lambda: ops.warp_rev_bundle_u8(col_d, Hc_d, 4, 4), 200, flush_l2=False)
    col_np = col_h.numpy()[0]
    def cpu_wrb_fn():
        parts = []
        for i in range(4):
            row = []
            for j in range(4):
                tmp = cv2.warpPerspective(col_np, Hc_np[i, j], dsize=(wn, hn), flags=cv2.WARP_INVERSE_MAP | cv2.INTER_LINEAR)
                row.append(tmp[i * 72:(i + 1) * 72, j * 128:(j + 1) * 128])
            parts.append(np.concatenate(row, 1))
        return np.concatenate(parts, 0)
    want = cpu_wrb_fn()
    assert np.array_equal(ops.warp_rev_bundle_u8(col_d, Hc_d, 4, 4).cpu().numpy()[0], want)
    tt = []
    for _ in range(20):
        t0 = time.perf_counter(); cpu_wrb_fn(); tt.append((time.perf_counter() - t0) * 1e6)
    cpu_wrb = float(np.median(tt))
except Exception as e:      # noqa: BLE001
    cpu_wrb = str(e)
# the frame producer, cvt_img2train (config.py:6-21): 1080p BGR uint8 -> gray -> Pillow bilinear to 288x512 -> fp32
dev_cvt = cpu_cvt = None
try:
    import cv2
    from PIL import Image
    bgr_np = np.random.RandomState(8).randint(0, 256, (1080, 1920, 3)).astype(np.uint8)
    bgr_d = torch.as_tensor(bgr_np).cuda()
    mgw.deploy.cvt_img2train(bgr_d)
    with torch.cuda.stream(s):
        dev_cvt = dev_time(lambda: mgw.deploy.cvt_img2train(bgr_d), 200, flush_l2=False)
    def cpu_cvt_fn():
        im = Image.fromarray(cv2.cvtColor(bgr_np, cv2.COLOR_BGR2GRAY)).resize((wn, hn), Image.BILINEAR)
        return (np.array(im) * (1. / 255) - 0.5).reshape((1, hn, wn, 1))
    assert np.array_equal(mgw.deploy.cvt_img2train(bgr_d).cpu().numpy(), cpu_cvt_fn().astype(np.float32))
    tt = []
    for _ in range(30):
        t0 = time.perf_counter(); cpu_cvt_fn(); tt.append((time.perf_counter() - t0) * 1e6)
    cpu_cvt = float(np.median(tt))
except Exception as e:      # noqa: BLE001
    cpu_cvt = str(e)
# the whole per-frame deploy loop on the device (deploy_bundle.py:259-328 minus the network, whose output is a fixed mesh here):
# H2D of the raw 1080p BGR frame -> cvt_img2train -> input assembly from the rings -> K1 + K2 on the gray frame -> ring push +
# black accumulation -> cv2.resize of the colour frame -> warpRevBundle2 -> D2H of the stabilised 512x288 colour frame
loop_p50 = loop_p99 = loop_graph_p50 = loop_graph_p99 = None
try:
    raw_h = torch.as_tensor(np.random.RandomState(9).randint(0, 256, (1080, 1920, 3)).astype(np.uint8)).pin_memory()
    raw_d = torch.empty((1080, 1920, 3), device=dev, dtype=torch.uint8)
    out_hh = torch.empty((hn, wn, 3), dtype=torch.uint8).pin_memory()
    st_loop = mgw.StreamState(gray_h[0, ..., 0])
    crop = mgw.CropState(hn, wn)
    with torch.cuda.stream(s):
        def loop_frame():
            raw_d.copy_(raw_h, non_blocking=True)
            cur = mgw.deploy.cvt_img2train(raw_d)                               # [1,288,512,1] fp32
            in_x = st_loop.assemble(cur.reshape(hn, wn))                        # the network's input (unused: no network here)
            out, black, img = mgw.transformer(cur, th)
            st_loop.push(out.reshape(hn, wn), black.reshape(hn, wn))
            crop.add(black)
            small = mgw.deploy.cv2_resize(raw_d, (wn, hn))
            dst = ops.remap_bundle_u8(small.reshape(1, hn, wn, 3), img)
            out_hh.copy_(dst[0], non_blocking=True)
            s.synchronize()
        for _ in range(20):
            loop_frame()
        lat = []
        for _ in range(500):
            t0 = time.perf_counter(); loop_frame(); lat.append((time.perf_counter() - t0) * 1e6)
    loop_p50, loop_p99 = float(np.percentile(lat, 50)), float(np.percentile(lat, 99))
    # the same loop captured ONCE in a CUDA graph (ring head on the device, static buffers, the two PCIe copies inside it)
    st_g = mgw.StreamState(gray_h[0, ..., 0], device_head=True)
    in_x_g = torch.empty((1, hn, wn, 13), device=dev)
    with torch.cuda.stream(s):
        def graph_frame():
            raw_d.copy_(raw_h, non_blocking=True)
            cur = mgw.deploy.cvt_img2train(raw_d)
            st_g.assemble(cur.reshape(hn, wn), out=in_x_g)
            out, black, img = mgw.transformer(cur, th)
            st_g.push(out.reshape(hn, wn), black.reshape(hn, wn))
            crop.add(black)
            small = mgw.deploy.cv2_resize(raw_d, (wn, hn))
            dst = ops.remap_bundle_u8(small.reshape(1, hn, wn, 3), img)
            out_hh.copy_(dst[0], non_blocking=True)
        for _ in range(3):
            graph_frame()
        s.synchronize()
        want_frame = out_hh.clone()
        gl = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gl, stream=s):
            graph_frame()
        latg = []
        for _ in range(500):
            t0 = time.perf_counter(); gl.replay(); s.synchronize(); latg.append((time.perf_counter() - t0) * 1e6)
        assert torch.equal(out_hh, want_frame)
    loop_graph_p50, loop_graph_p99 = float(np.percentile(latg, 50)), float(np.percentile(latg, 99))
except Exception as e:      # noqa: BLE001
    loop_p50 = str(e)
# the streaming state around it (deploy_bundle.py:259-274,319-328): input assembly from the history rings + push of the new frame
state = mgw.StreamState(gray_h[0, ..., 0])
cur2d = gray_d[0, ..., 0].contiguous()
img2d, blk2d = o1[0, ..., 0].contiguous(), b1[0].contiguous()
with torch.cuda.stream(s):
    def state_call():
        in_x = state.assemble(cur2d)
        state.push(img2d, blk2d)
        return in_x
    dev_state = dev_time(state_call, 200, flush_l2=False)
cpu_state_us = None
try:
    import collections
    fr = [np.zeros((1, hn, wn, 1), np.float32) for _ in range(32)]; mk = [np.zeros((1, hn, wn, 1), np.float32) for _ in range(32)]
    curn = gray_h.numpy().reshape(1, hn, wn, 1); imgn = np.zeros((hn, wn), np.float32); blkn = np.zeros((hn, wn), np.float32)
    def cpu_state():
        taps = (1, 2, 4, 8, 16, 32)
        x = np.concatenate([mk[-i] for i in taps] + [fr[-i] for i in taps] + [curn], axis=3)
        fr.append((imgn + blkn * (-1)).reshape(1, hn, wn, 1)); mk.append(blkn.reshape(1, hn, wn, 1)); fr.pop(0); mk.pop(0)
        return x
    tt = []
    for _ in range(50):
        t0 = time.perf_counter(); cpu_state(); tt.append((time.perf_counter() - t0) * 1e6)
    cpu_state_us = float(np.median(tt))
except Exception as e:      # noqa: BLE001
    cpu_state_us = str(e)
cpu_us = None
try:
    import cv2
    xm = xy1[0, ..., 0].cpu().numpy().copy(); ym = xy1[0, ..., 1].cpu().numpy().copy(); cimg = col_h[0].numpy()
    def cpu_remap():
        a = cv2.resize(cv2.resize(xm, (wn // 4, hn // 4)), (wn, hn)); b = cv2.resize(cv2.resize(ym, (wn // 4, hn // 4)), (wn, hn))
        return cv2.remap(cimg, (a + 1) / 2 * wn, (b + 1) / 2 * hn, cv2.INTER_LINEAR)
    cpu_remap()
    tt = []
    for _ in range(50):
        t0 = time.perf_counter(); cpu_remap(); tt.append((time.perf_counter() - t0) * 1e6)
    cpu_us = float(np.median(tt))
except Exception as e:      # noqa: BLE001
    cpu_us = 'cv2 unavailable: %s' % e
res['deploy_frame_288x512'] = {
    'us_device_graph_replay_warp_plus_remap': dev_frame, 'us_device_remap_only': dev_remap,
    'us_p50_host_to_host_u8_frames': float(np.percentile(lat3, 50)), 'us_p99_host_to_host': float(np.percentile(lat3, 99)),
    'us_cpu_opencv_remap_only': cpu_us, 'cpu_threads': os.cpu_count(),
    'us_p50_whole_deploy_loop_1080p_bgr_in_288x512_bgr_out_host_to_host': loop_p50, 'us_p99_whole_deploy_loop': loop_p99,
    'us_p50_whole_deploy_loop_one_cuda_graph': loop_graph_p50, 'us_p99_whole_deploy_loop_one_cuda_graph': loop_graph_p99,
    'us_device_cvt_img2train_1080p': dev_cvt, 'us_cpu_cv2_pil_cvt_img2train_1080p': cpu_cvt,
    'us_device_warp_rev_bundle': dev_wrb, 'us_cpu_opencv_warp_rev_bundle': cpu_wrb,
    'us_device_stream_state_assemble_plus_push': dev_state, 'us_cpu_numpy_stream_state': cpu_state_us,
    'note': '1x288x512: H2D gray fp32 + colour u8, K1 + K2 (C=1) + maps/4 + remap, D2H colour u8; CPU = the same three cv2 calls'}


# ---- config #3: StabNet forward, batch 16, 13-channel 288x512 input (configs/v2_93.py:19-22,40): the package's carrier
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = mgw.StabNet(in_ch=13, grid=(4, 4)).to(dev).eval().to(memory_format=torch.channels_last)
n = 16
x_nhwc = torch.tensor(synth.noise_image(n, 288, 512, 13, 5), device=dev)
with torch.no_grad():
    def backbone():
        return net(x_nhwc)

    def warp_stage(head):
        _, pts2 = mgw.get_4_pts(head, n, (4, 4))
        cur = x_nhwc[..., 12:13].contiguous()             # the current frame, channel 12 (s_net_bundle_nobm.py:281)
        return mgw.transformer(cur, pts2)

    head = backbone()
    t_backbone = dev_time(backbone, 20)
    t_warp = dev_time(lambda: warp_stage(head), 50)
    t_total = dev_time(lambda: warp_stage(backbone()), 20)
res['config3_stabnet_fwd_b16'] = {'us_backbone_torch_fp32': t_backbone, 'us_warp_stage_C1': t_warp, 'us_total': t_total,
                                  'warp_share_pct': 100 * t_warp / t_total,
                                  'note': 'backbone = dovs_b200.StabNet (torch/cuDNN ResNet-50-v2 carrier, random init; not a kernel-writing '
                                          'target); warp stage = get_4_pts + slice + K1 + K2 on the 1-channel current frame'}

# ---- config #5, one rank's share: the training objective of train_bundle_nobm.py:107-141 at 32 clips per GPU (256 over 8)
net.train()
nb = 32


def clip_batch(seed):
    return dict(x=torch.tensor(synth.noise_image(nb, 288, 512, 13, seed), device=dev),
                y=torch.tensor(synth.noise_image(nb, 288, 512, 1, seed + 1), device=dev),
                matches=torch.tensor(synth.uniform((nb, 3000, 4), -1, 1, seed + 2), device=dev),
                mask=(torch.rand(nb, 3000, device=dev) < 0.3).float())


b1, b2 = clip_batch(40), clip_batch(50)
# optical flow between the two clips as sampling coordinates: identity grid + a smooth displacement of a few pixels (a flow of
# independent uniform samples would make every gather and every atomic of temp_loss a random access, which no video does)
_ys, _xs = torch.meshgrid(torch.linspace(-1, 1, 288, device=dev), torch.linspace(-1, 1, 512, device=dev), indexing='ij')
flow = torch.stack([_xs + 0.03 * torch.sin(3 * _ys), _ys + 0.03 * torch.cos(2 * _xs)], -1).expand(nb, 288, 512, 2).contiguous()
gates = mgw.loss_gates(6000)
opt = torch.optim.Adam(net.parameters(), lr=2e-5)


def train_step():
    opt.zero_grad(set_to_none=True)
    total, _, _, _ = mgw.train_losses(net, b1, b2, flow, gates, batch_size=256)
    total.backward()
    opt.step()
    return total


l0 = mgw.launch_count()
train_step()
ours = mgw.launch_count() - l0
t_step = dev_time(train_step, 10, flush_l2=False)


def path_only():
    # the same step with the backbone cut out: theta is a leaf
    th = theta_leaf.detach().requires_grad_(True)
    tot = 0
    rets = []
    for b in (b1, b2):
        x = b['x'][..., 12:13].contiguous()
        if FUSED_PASS:      # everything after the head as one autograd node (mgw_train_pass_fwd / _bwd)
            t, _, out, black, _, _, _ = mgw.train_pass(th, x, b['y'], b['matches'], b['mask'], batch_size=256)
        else:               # the separate operators composed by torch autograd
            p1, p2 = mgw.get_4_pts(th, grid=(4, 4))
            il, out, black, fl = mgw.transformer_img_loss(x, p2, b['y'], batch_size=256)
            ftl, _ = mgw.feature_loss(b['matches'], b['mask'], fl, batch_size=256)
            t, _ = mgw.total_loss(th, p1, p2, il, ftl, batch_size=256)
        tot = tot + t
        rets.append((out, black))
    tot = tot + 500.0 * mgw.temp_loss(rets[0][0], rets[0][1], rets[1][0], rets[1][1], flow, batch_size=256)
    tot.backward()
    return tot


with torch.no_grad():
    theta_leaf = net(b1['x'])
path = {}
for FUSED_PASS in (False, True):
    l0 = mgw.launch_count()
    path_only()
    nl = mgw.launch_count() - l0
    t_path = dev_time(path_only, 10, flush_l2=False)
    # the same objective captured once in a CUDA graph (no host synchronisation anywhere in it: the loss backwards read their
    # upstream gradient on the device), replayed
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            path_only()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        path_only()
    path['fused' if FUSED_PASS else 'composed'] = dict(us_eager=t_path, us_graph_replay=dev_time(graph.replay, 20, flush_l2=False),
                                                      launches_of_this_library=nl)
t_path, t_path_graph = path['fused']['us_eager'], path['fused']['us_graph_replay']
res['config5_train_step_32_clips_per_gpu'] = {
    'us_step_backbone_plus_path_plus_adam': t_step, 'us_path_only_fwd_bwd_all_losses_eager': t_path, 'us_path_only_graph_replay': t_path_graph,
    'path_share_pct_eager': 100 * t_path / t_step, 'path_only': path,
    'launches_of_this_library_per_step': ours,
    'note': 'two passes (shared weights) + temp_loss, every loss term of the reference objective, Adam; backbone = torch fp32 carrier; '
            'the path = get_4_pts + fused warp/img_loss fwd+bwd + feature/temp/vertex losses on the 1-channel current frame; fused = one '
            'autograd node per pass (mgw_train_pass_fwd/bwd), composed = the separate operators under torch autograd'}
print(json.dumps(res, indent=1))
