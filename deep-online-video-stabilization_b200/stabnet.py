"""StabNet around the warp path: the backbone + regression head as a plain torch module (cuDNN / cuBLAS carrier, not a
kernel-writing target -- SURVEY.md 8(f) rank 3) wired to this library's vertex builder, multi-grid warp and losses.

Mirrors reference s_net_bundle_nobm.py:
  get_resnet            :250-264   slim resnet_v2_50(global_pool=False, output_stride=32) -> mean over H,W -> fc 2048 -> 1024
                                   -> 512 (ReLU) -> output_layer((grid_h+1)*(grid_w+1)*2)  (resnet.py:44-56)
  inference_stable_net  :266-381   x = the current frame (channel 2*before_ch of x_tensor with input_mask), theta -> get_4_pts ->
                                   transformer(x, pts2), the loss terms and total_loss, returned as the reference's `ret` dict
  train step            train_bundle_nobm.py:107-141   two passes (shared weights) + temp_loss

PARITY UNPINNED for the backbone itself: tf.contrib.slim is not in this image and the reference ships no checkpoint or
activations, so the ResNet is a structural restatement of slim's resnet_v2_50 (pre-activation bottlenecks 3-4-6-3, stride in
the last unit of blocks 1-3, 'SAME' padding, BN eps 1e-5) with torch's default initialisers.  Everything downstream of theta
IS pinned (tests/golden): get_4_pts, the warp, every loss term.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import losses
from .spatial_transformer3 import transformer


def _conv_same(cin, cout, k, stride):
    # slim conv2d_same: explicit symmetric padding (k-1)//2 for strided convs, 'SAME' otherwise -- identical for odd k
    return nn.Conv2d(cin, cout, k, stride, padding=(k - 1) // 2, bias=False)


class BottleneckV2(nn.Module):
    """slim resnet_v2.bottleneck: preact BN+ReLU, 1x1 -> 3x3 (stride) -> 1x1, shortcut = 1x1 conv of the preact (when the depth
    changes) or a stride-subsampled identity."""

    def __init__(self, cin, depth, stride):
        super().__init__()
        mid = depth // 4
        self.preact = nn.BatchNorm2d(cin, eps=1e-5, momentum=0.003)
        self.shortcut = nn.Conv2d(cin, depth, 1, stride, bias=True) if cin != depth else None
        self.stride = stride
        self.conv1, self.bn1 = _conv_same(cin, mid, 1, 1), nn.BatchNorm2d(mid, eps=1e-5, momentum=0.003)
        self.conv2, self.bn2 = _conv_same(mid, mid, 3, stride), nn.BatchNorm2d(mid, eps=1e-5, momentum=0.003)
        self.conv3 = nn.Conv2d(mid, depth, 1, bias=True)

    def forward(self, x):
        pre = TF.relu(self.preact(x))
        if self.shortcut is not None:
            sc = self.shortcut(pre)
        else:
            sc = x if self.stride == 1 else TF.max_pool2d(x, 1, self.stride)
        y = TF.relu(self.bn1(self.conv1(pre)))
        y = TF.relu(self.bn2(self.conv2(y)))
        return sc + self.conv3(y)


class ResNetV2_50(nn.Module):
    def __init__(self, in_ch):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, 64, 7, 2, padding=3, bias=True)
        units, cin = [], 64
        for depth, n, stride in ((256, 3, 2), (512, 4, 2), (1024, 6, 2), (2048, 3, 1)):
            for u in range(n):
                units.append(BottleneckV2(cin, depth, stride if u == n - 1 else 1))
                cin = depth
        self.units = nn.Sequential(*units)
        self.postnorm = nn.BatchNorm2d(2048, eps=1e-5, momentum=0.003)

    def forward(self, x):
        x = self.conv1(x)
        x = TF.max_pool2d(TF.pad(x, (0, 1, 0, 1), value=float('-inf')), 3, 2)       # 'SAME' 3x3/2 on an even size pads after
        return TF.relu(self.postnorm(self.units(x)))


class StabNet(nn.Module):
    """x_tensor [N,H,W,Cin] (NHWC like the reference's placeholder, :278) -> theta [N, 2*(gh+1)*(gw+1)]."""

    def __init__(self, in_ch=13, grid=(4, 4)):
        super().__init__()
        self.grid = (int(grid[0]), int(grid[1]))
        self.backbone = ResNetV2_50(in_ch)
        self.fc = nn.Sequential(nn.Linear(2048, 2048), nn.ReLU(), nn.Linear(2048, 1024), nn.ReLU(), nn.Linear(1024, 512), nn.ReLU())
        n_out = (self.grid[0] + 1) * (self.grid[1] + 1) * 2
        self.head = nn.Linear(512, n_out)
        # output_layer (resnet.py:50-53): uniform_unit_scaling_initializer(factor=1.0) = U(+-sqrt(3 / fan_in)), zero bias
        lim = math.sqrt(3.0 / 512)
        nn.init.uniform_(self.head.weight, -lim, lim)
        nn.init.zeros_(self.head.bias)

    def features(self, x_tensor):
        x = x_tensor.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        return self.fc(self.backbone(x).mean(dim=(2, 3)))            # tf.reduce_mean(resnet, [1, 2])   (:254)

    def forward(self, x_tensor):
        return self.head(self.features(x_tensor))


class _Ret(dict):
    """the reference's `ret`; 'error' = |h_trans - y| (a display tensor, :361) is formed when somebody asks for it"""

    def __missing__(self, key):
        if key != 'error':
            raise KeyError(key)
        self[key] = (self['output'] - self['y']).abs()
        return self[key]


def inference_stable_net(net, x_tensor, y, matches, mask, use_black_loss=1.0, use_theta_only=0.0, before_ch=6, input_mask=True,
                         mul=None, do_crop_rate=0.8, batch_size=None, regu_loss=0.0, fused=True):
    """One pass of reference s_net_bundle_nobm.py:266-381 -> the `ret` dict (same keys; 'theta', 'pts2', 'flow' added).
    batch_size: the divisor of the batch-mean terms (GLOBAL batch under data parallelism).  fused: everything after the head as
    one autograd node (losses.train_pass); False composes the separate operators (same values)."""
    cur = 2 * before_ch if input_mask else before_ch                      # :282-285
    x = x_tensor[..., cur:cur + 1].contiguous()
    theta = net(x_tensor)
    if fused:
        total, parts, h_trans, black_pix, flow, _, pts2 = losses.train_pass(
            theta, x, y, matches, mask, regu_loss=regu_loss, use_black_loss=use_black_loss, use_theta_only=use_theta_only, mul=mul,
            grid=net.grid, do_crop_rate=do_crop_rate, batch_size=batch_size)
        n, h, w, _ = h_trans.shape
        ret = _Ret(parts)
        ret.update(black_pix=black_pix.reshape(n, h, w, 1), mask=mask, matches=matches, x_tensor=x_tensor, use_theta_only=use_theta_only,
                   y=y, output=h_trans, total_loss=total, theta=theta, pts2=pts2, flow=flow)
        return ret
    pts1, pts2 = losses.get_4_pts(theta, grid=net.grid, do_crop_rate=do_crop_rate)
    img_l, h_trans, black_pix, flow = losses.transformer_img_loss(x, pts2, y, batch_size=batch_size)       # :332,:347-352
    feat_l, _ = losses.feature_loss(matches, mask, flow, batch_size=batch_size)                             # :335-343
    total, parts = losses.total_loss(theta, pts1, pts2, img_l, feat_l, regu_loss=regu_loss, use_black_loss=use_black_loss,
                                     use_theta_only=use_theta_only, mul=mul, do_crop_rate=do_crop_rate, batch_size=batch_size)
    n, h, w, _ = h_trans.shape
    ret = _Ret(parts)
    ret.update(black_pix=black_pix.reshape(n, h, w, 1), mask=mask, matches=matches, x_tensor=x_tensor,
               use_theta_only=use_theta_only, y=y, output=h_trans, total_loss=total, theta=theta, pts2=pts2, flow=flow)
    return ret


def train_losses(net, batch1, batch2, flow, gates, mul=None, batch_size=None, fused=True):
    """The training objective of train_bundle_nobm.py:107-141: two passes with shared weights and the temporal loss between
    them.  batch1/2 = dict(x, y, matches, mask); gates = losses.loss_gates(i).  -> (total_loss, ret1, ret2, temp_loss)"""
    m = dict(losses.V2_93_MULS)
    m.update(mul or {})
    kw = dict(use_black_loss=float(gates['use_black']), use_theta_only=float(gates['theta_only']), mul=m, batch_size=batch_size, fused=fused)
    ret1 = inference_stable_net(net, batch1['x'], batch1['y'], batch1['matches'], batch1['mask'], **kw)
    ret2 = inference_stable_net(net, batch2['x'], batch2['y'], batch2['matches'], batch2['mask'], **kw)
    t_loss = losses.temp_loss(ret1['output'], ret1['black_pix'], ret2['output'], ret2['black_pix'], flow,
                              use_temp_loss=float(gates['use_temp']), batch_size=batch_size)
    return ret1['total_loss'] + ret2['total_loss'] + t_loss * m['temp_mul'], ret1, ret2, t_loss
