"""Data parallelism for the warp path: one process per GPU, frames/clips sharded along the batch, NO communication
in the forward pass; the only collective is an all-reduce of the mesh-head gradient.

The reference is single-GPU (no collective anywhere).  Under data parallelism every sample is independent through
the whole path (theta is a per-sample activation), so the only cross-rank quantity downstream of the path is the
gradient of the regression head that produces theta: fc_weights [512, 2*(gh+1)*(gw+1)] + bias (reference
resnet.py:51-53 via s_net_bundle_nobm.py:259) = 25 650 fp32 = 100.2 KB for the 4x4 mesh -- latency-bound, so it is
issued on a side stream and overlapped with the next step's forward.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """(rank, world, local_rank); initialises torch.distributed when WORLD_SIZE > 1 (torchrun env)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(n_global, rank, world):
    """contiguous batch slice [lo, hi) of rank `rank`; the first n_global % world ranks take one extra sample."""
    base, extra = divmod(n_global, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t, rank, world):
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi].contiguous()


class MeshHeadGradReducer:
    """Sums the mesh-head gradient over ranks with one all-reduce per step on a side stream.

    head_grad(features [n_local,F], dtheta [n_local,gh+1,gw+1,2], slot) -> flat [F*V + V] buffer (dW = features^T . dtheta,
    db = sum_n dtheta), already scaled for the GLOBAL batch by the losses.  launch(slot) forks the side stream from the
    current one and starts the all-reduce of that slot's buffer there; wait() joins it back (call it right before the
    optimizer step that consumes the buffer).  With nbuf > 1 a step can compute its gradient into one slot while the previous
    step's slot is still being reduced -- the form that can be captured in a CUDA graph (fork and join inside one capture).
    wait() must come before anything rewrites bufs[slot] (the reduction owns the slot until then).
    """

    def __init__(self, n_features, n_out, device, nbuf=1):
        self.bufs = [torch.zeros(n_features * n_out + n_out, device=device, dtype=torch.float32) for _ in range(nbuf)]
        self.nf, self.no = n_features, n_out
        # high priority: the few CTAs of the reduction are placed ahead of the queued CTAs of the warp kernels it overlaps
        self.side = torch.cuda.Stream(device=device, priority=-1) if torch.device(device).type == 'cuda' else None
        self.work = None
        self.last = 0

    @property
    def buf(self):
        return self.bufs[0]

    def head_grad(self, features, dtheta, slot=0):
        buf = self.bufs[slot]
        d = dtheta.reshape(dtheta.shape[0], self.no)
        torch.mm(features.t(), d, out=buf[:self.nf * self.no].view(self.nf, self.no))
        torch.sum(d, dim=0, out=buf[self.nf * self.no:])
        return buf

    def launch(self, slot=0, features=None, dtheta=None):
        """starts the all-reduce of slot `slot` on the side stream; with (features, dtheta) the head gradient itself is
        formed there first, so nothing of the reduction sits on the stream that runs the warp kernels."""
        self.last = slot
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.side is None:
            if features is not None:
                self.head_grad(features, dtheta, slot)
            if multi:
                self.work = dist.all_reduce(self.bufs[slot], op=dist.ReduceOp.SUM, async_op=True)
            return
        if not multi and features is None:
            return
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            if features is not None:
                # the side stream reads tensors allocated on the caller's stream: tell the caching allocator, or a backward
                # temporary dropped right after launch() could be handed out again while the GEMM still reads it
                features.record_stream(self.side)
                dtheta.record_stream(self.side)
                self.head_grad(features, dtheta, slot)
            if multi:
                dist.all_reduce(self.bufs[slot], op=dist.ReduceOp.SUM)

    def wait(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        elif self.work is not None:
            self.work.wait()
            self.work = None
        return self.bufs[self.last]


def allreduce_grads(params, bucket_bytes=32 << 20):
    """Data-parallel training of the whole network (the carrier of SURVEY.md 8(f) rank 3): sums every parameter gradient over
    the ranks in flat fp32 buckets (losses are already divided by the GLOBAL batch, so SUM is the reduction).  No-op for one rank."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    for g in grads:
        bucket.append(g)
        size += g.numel() * 4
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()
