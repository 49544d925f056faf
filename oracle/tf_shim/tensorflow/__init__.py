"""TEST INFRASTRUCTURE -- not product code.

A minimal stand-in for the `tensorflow` (1.x graph API) module, backed by eager
torch-CPU tensors, whose only purpose is to let the *unmodified* reference
sources under /root/reference (spatial_transformer3.py, spatial_transformer.py
and selected functions of s_net_bundle_nobm.py / train_bundle_nobm.py) execute
in this container, where TensorFlow itself is not installed.  It is put on
sys.path by oracle/make_golden.py and by the `reference`-marked CPU tests, and
is never imported by the product package.

Semantics that matter for numerics:
  * linspace is fp32 `start + i*step`, step = (stop-start)/(n-1) in fp32
    (TF 1.x LinSpaceOp), NOT torch.linspace's symmetric formula.
  * add_n sums left to right.
  * round is half-to-even (tf.round).
  * floor/cast/clip on integers cut autograd exactly where TF does.
  * tensors created with name=... are recorded in NAMED (deploy_bundle.py
    fetches 'Hs', 'x_map', 'y_map', 'black_pix', 'output_img' by name).
"""
import contextlib

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64

NAMED = {}          # name -> last tensor registered under that name
DEFAULT_FLOAT = [torch.float32]   # oracle/make_golden.py flips this to float64 for the arbiter run


def _dt(d):
    if d is None:
        return DEFAULT_FLOAT[0]
    if isinstance(d, str):
        d = {'float32': torch.float32, 'int32': torch.int32, 'float64': torch.float64,
             'int64': torch.int64}[d]
    if d is torch.float32:
        return DEFAULT_FLOAT[0]
    return d


def _reg(t, name):
    if name is not None:
        NAMED[name] = t
    return t


def _int(v):
    if isinstance(v, torch.Tensor):
        return int(v.item())
    return int(v)


def _ints(seq):
    if isinstance(seq, torch.Tensor):
        return [int(v) for v in seq.reshape(-1).tolist()]
    if isinstance(seq, (int, np.integer)):
        return [int(seq)]
    return [_int(v) for v in seq]


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (float, np.floating)):
        return torch.tensor(x, dtype=DEFAULT_FLOAT[0])
    if isinstance(x, (int, np.integer)):
        return torch.tensor(x, dtype=torch.int32)
    arr = np.asarray(x)
    if arr.dtype.kind == 'f':
        return torch.tensor(arr, dtype=DEFAULT_FLOAT[0])
    return torch.tensor(arr)


@contextlib.contextmanager
def variable_scope(*a, **k):
    yield


@contextlib.contextmanager
def name_scope(*a, **k):
    yield


def shape(t):
    return [int(s) for s in t.shape]


def cast(x, dtype):
    return _t(x).to(_dt(dtype))


def reshape(t, shp, name=None):
    return _reg(_t(t).reshape(_ints(shp)), name)


def slice(t, begin, size):           # noqa: A001 - mirrors tf.slice
    idx = []
    for b, s in zip(_ints(begin), _ints(size)):
        idx.append(np.s_[b:] if s == -1 else np.s_[b:b + s])
    return t[tuple(idx)]


def concat(values, axis=0, name=None):
    return _reg(torch.cat([_t(v) for v in values], dim=axis), name)


def stack(values, axis=0):
    if all(not isinstance(v, torch.Tensor) or v.dim() == 0 for v in values) and \
            all(not (isinstance(v, torch.Tensor) and v.is_floating_point()) and
                not isinstance(v, float) for v in values):
        return torch.tensor([_int(v) for v in values], dtype=torch.int32)
    return torch.stack([_t(v) for v in values], dim=axis)


def tile(t, multiples):
    return t.repeat(*_ints(multiples))


def ones(shape, dtype=None):          # noqa: A002
    return torch.ones(_ints(shape), dtype=_dt(dtype))


def zeros(shape, dtype=None):         # noqa: A002
    return torch.zeros(_ints(shape), dtype=_dt(dtype))


def ones_like(t):
    return torch.ones_like(t)


def constant(value, shape=None, dtype=None):   # noqa: A002
    if _dt(dtype).is_floating_point:
        # a tf.float32 constant is an fp32 INPUT of the graph: in the fp64 arbiter run it keeps its fp32-rounded value
        arr = np.asarray(value, dtype=np.float32 if dtype in (torch.float32, 'float32') else np.float64).astype(np.float64)
    else:
        arr = np.asarray(value)
    t = torch.tensor(arr, dtype=_dt(dtype))
    if shape is not None:
        t = t.reshape(_ints(shape))
    return t


def eye(n):
    return torch.eye(n, dtype=DEFAULT_FLOAT[0])


def expand_dims(t, axis):
    return t.unsqueeze(axis)


def transpose(t, perm):
    return t.permute(*perm)


def matmul(a, b):
    if not a.is_floating_point():
        return (a.to(torch.int64) @ b.to(torch.int64)).to(a.dtype)
    return a @ b


def matrix_inverse(a):
    return torch.linalg.inv(a)


def linspace(start, stop, num):
    dt = DEFAULT_FLOAT[0]
    start_t = torch.tensor(start, dtype=dt)
    step = (torch.tensor(stop, dtype=dt) - start_t) / torch.tensor(num - 1, dtype=dt)
    return start_t + step * torch.arange(num, dtype=dt)


def range(n):                          # noqa: A001
    return torch.arange(_int(n), dtype=torch.int32)


def floor(x):
    return torch.floor(x)


def round(x):                          # noqa: A001
    return torch.round(x)              # half-to-even, like tf.round


def abs(x):                            # noqa: A001
    return torch.abs(x)


def clip_by_value(x, lo, hi):
    lo = lo.item() if isinstance(lo, torch.Tensor) else lo
    hi = hi.item() if isinstance(hi, torch.Tensor) else hi
    return torch.clamp(x, lo, hi)


def maximum(a, b):
    return torch.maximum(_t(a), _t(b).to(_t(a).dtype))


def minimum(a, b):
    return torch.minimum(_t(a), _t(b).to(_t(a).dtype))


def greater(a, b):
    return a > b


def logical_or(a, b):
    return a | b


def where(cond, a, b):
    return torch.where(cond, a, b)


def gather(params, indices):
    return params[indices.to(torch.int64)]


def add_n(xs):
    out = xs[0]
    for x in xs[1:]:
        out = out + x
    return out


def div(a, b, name=None):
    return _reg(a / b, name)


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(axis)
    return (axis,)


def reduce_sum(x, axis=None, **kw):
    ax = _axes(axis)
    return x.sum() if ax is None else x.sum(dim=ax)


def reduce_mean(x, axis=None, **kw):
    ax = _axes(axis)
    return x.mean() if ax is None else x.mean(dim=ax)


def placeholder(dtype, shape=None, name=None):   # noqa: A002
    raise RuntimeError('tf.placeholder: feed concrete tensors when driving the reference through the shim')


def Print(t, *a, **k):
    return t


def stop_gradient(t):
    return t.detach()


class _Summary:
    @staticmethod
    def tensor_summary(*a, **k):
        return None

    @staticmethod
    def image(*a, **k):
        return None

    @staticmethod
    def scalar(*a, **k):
        return None


summary = _Summary()


def add_to_collection(*a, **k):
    return None
