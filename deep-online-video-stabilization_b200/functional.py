"""torch.autograd glue: each Function is one forward C-ABI call and one backward C-ABI call (see ops.py)."""
import torch

from . import ops


def empty_batch(U, theta, oh, ow):
    """N = 0 (an empty shard of a ragged batch split): nothing to launch; empty outputs that still hang off U and theta in
    the autograd graph, as the reference's graph would produce"""
    if not (U.is_cuda and theta.is_cuda):
        raise RuntimeError('U and theta must live on a CUDA device: this library has no CPU path')
    z = (U.sum() + theta.sum()) * 0
    return U.new_zeros((0, oh, ow, U.shape[3])) + z, U.new_zeros((0, oh, ow)), U.new_zeros((0, oh, ow, 2)) + z


class MeshWarp(torch.autograd.Function):
    """spatial_transformer3.transformer(U, theta) (reference spatial_transformer3.py:19-365)."""

    @staticmethod
    def forward(ctx, U, theta):
        out, black, img, Hs = ops.mesh_warp_fwd(U, theta)
        ctx.save_for_backward(U, theta, Hs)
        ctx.mark_non_differentiable(black, Hs)
        ctx.set_materialize_grads(False)          # an unused output arrives as None, not as a tensor of zeros to stream through the kernel
        return out, black, img, Hs

    @staticmethod
    def backward(ctx, d_out, _d_black, d_img, _d_Hs):
        U, theta, Hs = ctx.saved_tensors
        if d_out is None and d_img is None:
            return (torch.zeros_like(U) if ctx.needs_input_grad[0] else None), torch.zeros_like(theta)
        if d_out is None:
            d_out = torch.zeros_like(U)
        dU, dtheta = ops.mesh_warp_bwd(U, theta, Hs, d_out.contiguous(), None if d_img is None else d_img.contiguous(),
                                       want_dU=ctx.needs_input_grad[0])
        return dU, dtheta


class MeshWarpImgLoss(torch.autograd.Function):
    """transformer(U, theta) + img_loss(h_trans, y, black_pix) in one forward kernel and one backward kernel
    (reference s_net_bundle_nobm.py:332 and :347-352).  Returns (loss, output, black_pix, img).  The fused backward is
    taken when no other gradient reaches `output`; otherwise the loss gradient is materialised and added."""

    @staticmethod
    def forward(ctx, U, theta, y, batch):
        out, black, img, Hs, sums = ops.mesh_warp_img_loss_fwd(U, theta, y)
        ctx.save_for_backward(U, theta, Hs, out, y, black, sums)
        ctx.batch = batch
        ctx.mark_non_differentiable(black)
        ctx.set_materialize_grads(False)          # g_out is None unless something else consumes `output`: the fused backward's condition
        loss = ops.loss_ratio_sum(sums, 1.0 / batch)
        return loss, out, black, img

    @staticmethod
    def backward(ctx, g_loss, g_out, _g_black, g_img):
        U, theta, Hs, out, y, black, sums = ctx.saved_tensors
        g_img = None if g_img is None else g_img.contiguous()
        up = 0.0 if g_loss is None else g_loss                   # device scalar: read by the kernel, no host sync
        if g_loss is None and g_out is None and g_img is None:
            return (torch.zeros_like(U) if ctx.needs_input_grad[0] else None), torch.zeros_like(theta), None, None
        if g_out is None:
            dU, dtheta = ops.mesh_warp_img_loss_bwd(U, theta, Hs, out, y, black, sums, up, ctx.batch, g_img,
                                                    want_dU=ctx.needs_input_grad[0])
        else:
            # a second gradient on `output` (temp_loss): added to the loss gradient inside the warp backward
            dU, dtheta = ops.mesh_warp_img_loss_bwd(U, theta, Hs, out, y, black, sums, up, ctx.batch, g_img,
                                                    want_dU=ctx.needs_input_grad[0], d_out_extra=g_out.contiguous())
        return dU, dtheta, None, None


class TrainPass(torch.autograd.Function):
    """One training pass after the network head (reference s_net_bundle_nobm.py:266-381: get_4_pts, transformer, img_loss,
    feature_loss, the vertex regularisers, the weighted total) as 7 forward and 5 backward launches.
    -> (result [9] = total + weighted parts, h_trans, black_pix, flow, stable_warpped, pts2).  Differentiable outputs: result[0]
    through `total` and h_trans (another consumer of the warped frame, i.e. temp_loss, adds its gradient inside the warp
    backward); flow / pts2 / the parts are returned for display only."""

    _FWD_KEYS = ('pts1', 'pts2', 'Hs', 'out', 'black', 'img', 'acc')

    @staticmethod
    def forward(ctx, head, U, y, matches, mask, regu, coef, gh, gw, do_crop_rate):
        f = ops.train_pass_fwd(head, U, y, matches, mask, coef, gh, gw, do_crop_rate, regu=regu)
        # (outputs go through save_for_backward as well: kept on ctx directly they would form a reference cycle with the graph)
        ctx.save_for_backward(head, U, y, matches, mask, *[f[k] for k in TrainPass._FWD_KEYS])
        ctx.cfg = (coef, gh, gw, do_crop_rate)
        total, parts = f['result'][0], f['result'][1:]
        ctx.mark_non_differentiable(parts, f['black'], f['img'], f['warpped'], f['pts2'])
        ctx.set_materialize_grads(False)
        return total, parts, f['out'], f['black'], f['img'], f['warpped'], f['pts2']

    @staticmethod
    def backward(ctx, g_total, _g_parts, g_out, *_unused):
        head, U, y, matches, mask = ctx.saved_tensors[:5]
        fwd = dict(zip(TrainPass._FWD_KEYS, ctx.saved_tensors[5:]))
        coef, gh, gw, rate = ctx.cfg
        if g_total is None and g_out is None:
            return (torch.zeros_like(head), torch.zeros_like(U) if ctx.needs_input_grad[1] else None) + (None,) * 8
        if g_total is None:
            coef = (0.0,) * 9 + (coef[9], 0.0)      # only the gradient through h_trans is alive
        d_head, dU = ops.train_pass_bwd(head, U, y, matches, mask, fwd, coef, gh, gw, rate, g_total=g_total,
                                        d_out_extra=None if g_out is None else g_out.contiguous(), want_dU=ctx.needs_input_grad[1])
        regu_grad = None
        if ctx.needs_input_grad[5]:
            regu_grad = (g_total if g_total is not None else torch.zeros((), device=head.device)) * (coef[6] * coef[10])
        return d_head, dU, None, None, None, regu_grad, None, None, None, None


class HomographyWarp(torch.autograd.Function):
    """spatial_transformer.transformer(U, theta[N,9], out_size) (reference spatial_transformer.py:18-197)."""

    @staticmethod
    def forward(ctx, U, theta, out_size):
        out, black, _ = ops.homography_warp_fwd(U, theta, out_size)
        ctx.save_for_backward(U, theta)
        ctx.out_size = out_size
        ctx.mark_non_differentiable(black)
        return out, black

    @staticmethod
    def backward(ctx, d_out, _d_black):
        U, theta = ctx.saved_tensors
        dU, dtheta = ops.homography_warp_bwd(U, theta, d_out.contiguous(), ctx.out_size, want_dU=ctx.needs_input_grad[0])
        return dU, dtheta, None


class Interpolate(torch.autograd.Function):
    """interpolate(im, x, y, out_size) (reference spatial_transformer.py:200-281)."""

    @staticmethod
    def forward(ctx, im, x, y, out_size):
        out = ops.interp_fwd(im, x, y, out_size)
        ctx.save_for_backward(im, x, y)
        ctx.out_size = out_size
        return out

    @staticmethod
    def backward(ctx, d_out):
        im, x, y = ctx.saved_tensors
        need = ctx.needs_input_grad
        d_im, dx, dy = ops.interp_bwd(im, x, y, d_out.contiguous(), ctx.out_size, want_dim=need[0], want_dxy=need[1] or need[2])
        return d_im, dx, dy, None


class Vertices(torch.autograd.Function):
    """get_4_pts (reference s_net_bundle_nobm.py:29-71)."""

    @staticmethod
    def forward(ctx, head, gh, gw, do_crop_rate):
        pts1, pts2 = ops.vertices_fwd(head, gh, gw, do_crop_rate)
        ctx.save_for_backward(head)
        ctx.cfg = (gh, gw, do_crop_rate)
        return pts1, pts2

    @staticmethod
    def backward(ctx, d_pts1, d_pts2):
        (head,) = ctx.saved_tensors
        gh, gw, rate = ctx.cfg
        d_head = ops.vertices_bwd(head, None if d_pts2 is None else d_pts2.contiguous(),
                                  None if d_pts1 is None else d_pts1.contiguous(), gh, gw, rate)
        return d_head, None, None, None


class ImgLoss(torch.autograd.Function):
    """img_loss (reference s_net_bundle_nobm.py:347-352); batch = the divisor (global batch under data parallelism)."""

    @staticmethod
    def forward(ctx, out, y, black, batch):
        sums = ops.img_loss_fwd(out, y, black)
        ctx.save_for_backward(out, y, black, sums)
        ctx.batch = batch
        return ops.loss_ratio_sum(sums, 1.0 / batch)

    @staticmethod
    def backward(ctx, g):
        out, y, black, sums = ctx.saved_tensors
        d_out = ops.img_loss_bwd(out, y, black, sums, (out.shape[0] / ctx.batch, g))
        return d_out, (-d_out if ctx.needs_input_grad[1] else None), None, None


class FeatureLoss(torch.autograd.Function):
    """feature_loss with warp_pts (reference s_net_bundle_nobm.py:215-230,335-343)."""

    @staticmethod
    def forward(ctx, matches, mask, flow, batch):
        per, warpped = ops.feature_loss_fwd(matches, mask, flow)
        ctx.save_for_backward(matches, mask, flow)
        ctx.batch = batch
        ctx.mark_non_differentiable(warpped)
        return per.sum() / batch, warpped

    @staticmethod
    def backward(ctx, g, _gw):
        matches, mask, flow = ctx.saved_tensors
        d_flow = ops.feature_loss_bwd(matches, mask, flow, (flow.shape[0] / ctx.batch, g))
        return None, None, d_flow, None


class TempLoss(torch.autograd.Function):
    """temp_loss (reference train_bundle_nobm.py:115-125)."""

    @staticmethod
    def forward(ctx, out1, black1, out2, black2, flow, batch, use_temp_loss):
        sums = ops.temp_loss_fwd(out1, black1, out2, black2, flow)
        ctx.save_for_backward(out1, black1, out2, black2, flow, sums)
        ctx.cfg = (batch, use_temp_loss)
        return ops.loss_ratio_sum(sums, use_temp_loss / batch)

    @staticmethod
    def backward(ctx, g):
        out1, black1, out2, black2, flow, sums = ctx.saved_tensors
        batch, use = ctx.cfg
        d1, d2 = ops.temp_loss_bwd(out1, black1, out2, black2, flow, sums, (use * out1.shape[0] / batch, g))
        return d1, None, d2, None, None, None, None


class VertexLosses(torch.autograd.Function):
    """The four vertex regularisers (reference s_net_bundle_nobm.py:139-210,246-247) as SUMS [4]: |theta|, black_err^2,
    distortion residuals^2, lattice second differences^2.  Any input may be None."""

    @staticmethod
    def forward(ctx, theta, pts1, pts2, gh, gw, do_crop_rate):
        c = [None if t is None else t.contiguous() for t in (theta, pts1, pts2)]
        sums = ops.vertex_losses_fwd(c[0], c[1], c[2], gh, gw, do_crop_rate)
        ctx.present = [t is not None for t in c]
        ctx.save_for_backward(*[t for t in c if t is not None])
        ctx.cfg = (gh, gw, do_crop_rate)
        return sums

    @staticmethod
    def backward(ctx, g):
        it = iter(ctx.saved_tensors)
        theta, pts1, pts2 = [next(it) if p else None for p in ctx.present]
        gh, gw, rate = ctx.cfg
        need = tuple(ctx.needs_input_grad[k] for k in range(3))
        d = ops.vertex_losses_bwd(theta, pts1, pts2, gh, gw, g.contiguous(), rate, need)
        return d[0], d[1], d[2], None, None, None
