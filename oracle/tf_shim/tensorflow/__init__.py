"""Minimal torch-backed stand-in for the slice of the TensorFlow 2 API that the
reference's ``signals.py`` and ``model.py`` touch.  TEST INFRASTRUCTURE ONLY.

Purpose: TensorFlow cannot be installed offline, but the reference is plain Python
on top of ~50 TF ops.  Putting this directory first on ``sys.path`` lets the
reference's *own, unmodified source files* execute (float32 CPU torch tensors stand
in for tf.Tensors; ``tf.GradientTape`` maps to torch.autograd), which is how
``oracle/make_golden.py`` produces ``tests/golden/ref_shim_*.npz``.

Numerics deliberately mirrored from TensorFlow:
* ``tf.math.special.bessel_j0`` = Cephes single-precision j0f (Eigen generic_j0<float>),
  with the registered gradient ``-bessel_j1(x) * dy`` (Cephes j1f);
* ``tf.linspace`` / ``tf.range`` element formulas (start + i*delta in float32);
* Python scalars are converted to the tensor dtype before every op (torch does the same).
Everything else is the corresponding float32 torch op (<= 1 ulp from TF's Eigen kernels).

Random ops draw from torch's global generator and are *recorded* in ``random.LOG`` so
the oracle can be fed identical draws (TF's own Philox streams are not reproducible offline).
"""
import sys
import types

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
bool = torch.bool  # noqa: A001  (tf.bool)

dtypes = types.SimpleNamespace(float32=float32, float64=float64, int32=int32, int64=int64)

_builtin_bool = type(True)


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (torch.Size, tuple, list)) and all(isinstance(v, (int, np.integer)) for v in x):
        return torch.tensor(list(x), dtype=dtype or torch.int64)
    a = np.asarray(x)
    if dtype is None:
        dtype = torch.float32 if a.dtype.kind == 'f' else None
    return torch.as_tensor(a, dtype=dtype)


def _shape_arg(shape):
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.tolist()]
    return [int(v) for v in shape]


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


def constant(x, dtype=None):
    return _t(x, dtype)


def cast(x, dtype):
    return _t(x).to(dtype)


def shape(x):
    return _t(x).shape


def reshape(x, shp):
    return _t(x).reshape(_shape_arg(shp))


def split(x, num_or_sizes, axis=0):
    x = _t(x)
    if isinstance(num_or_sizes, int):
        return list(torch.split(x, x.shape[axis] // num_or_sizes, dim=axis))
    return list(torch.split(x, list(num_or_sizes), dim=axis))


def concat(values, axis):
    vals = [_t(v) for v in values]
    return torch.cat(vals, dim=axis)


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], dim=axis)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def squeeze(x, axis=None):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def zeros_like(x):
    return torch.zeros_like(_t(x))


def ones_like(x):
    return torch.ones_like(_t(x))


def zeros(shp, dtype=float32):
    return torch.zeros(_shape_arg(shp), dtype=dtype)


def ones(shp, dtype=float32):
    return torch.ones(_shape_arg(shp), dtype=dtype)


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    """tf.range: size = ceil(|limit-start|/|delta|), element i = start + i*delta in `dtype`."""
    if limit is None:
        start, limit = 0, start
    if dtype is None:
        dtype = float32 if any(isinstance(v, float) for v in (start, limit, delta)) else int32
    if dtype in (float32, float64):
        s, e, d = (torch.tensor(v, dtype=dtype) for v in (start, limit, delta))
        n = int(torch.ceil(torch.abs((e - s) / d)).item())
        return s + torch.arange(n, dtype=dtype) * d
    return torch.arange(start, limit, delta, dtype=dtype)


def linspace(start, stop, num):
    """tf.linspace (math_ops.linspace_nd): [start, start + delta*i (i=1..num-2), stop]."""
    start, stop = _t(start), _t(stop)
    delta = (stop - start) / torch.tensor(num - 1, dtype=start.dtype)
    mid = start + delta * torch.arange(1, num - 1, dtype=start.dtype)
    return torch.cat([start.reshape(1), mid, stop.reshape(1)])


def meshgrid(*args, indexing='xy'):
    return list(torch.meshgrid(*[_t(a) for a in args], indexing=indexing))


def square(x):
    x = _t(x)
    return x * x


def sqrt(x):
    return torch.sqrt(_t(x))


def exp(x):
    return torch.exp(_t(x))


def tanh(x):
    return torch.tanh(_t(x))


def abs(x):  # noqa: A001
    return torch.abs(_t(x))


def pow(x, y):  # noqa: A001
    return torch.pow(_t(x), _t(y))


def where(cond, a=None, b=None):
    if a is None:
        return torch.nonzero(cond)
    return torch.where(cond, _t(a), _t(b))


def clip_by_value(x, lo, hi):
    return torch.clamp(_t(x), lo, hi)


def stop_gradient(x):
    return _t(x).detach()


def logical_and(a, b):
    return torch.logical_and(a, b)


def _reduce(fn, x, axis, keepdims):
    x = _t(x)
    if axis is None:
        return fn(x)
    return fn(x, dim=axis, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    return _reduce(torch.sum, x, axis, keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    return _reduce(torch.mean, x, axis, keepdims)


def vectorized_map(fn, elems):
    """Per-element map over axis 0 (semantically what pfor computes)."""
    n = elems[0].shape[0]
    outs = [fn(tuple(e[i] for e in elems)) for i in _builtin_range(n)]
    return torch.stack(outs, 0)


import builtins as _b  # noqa: E402
_builtin_range = _b.range


# ---------------------------------------------------------------- Bessel (Cephes j0f / j1f)
def _polevl(x, coef):
    acc = torch.full_like(x, coef[0])
    for c in coef[1:]:
        acc = acc * x + c
    return acc


_JP = (-6.068350350393235e-8, 6.388945720783375e-6, -3.969646342510940e-4, 1.332913422519003e-2, -1.729150680240724e-1)
_MO = (-6.838999669318810e-2, 1.864949361379502e-1, -2.145007480346739e-1, 1.197549369473540e-1,
       -3.560281861530129e-3, -4.969382655296620e-2, -3.355424622293709e-6, 7.978845717621440e-1)
_PH = (3.242077816988247e1, -3.630592630518434e1, 1.756221482109099e1, -4.974978466280903e0,
       1.001973420681837e0, -1.939906941791308e-1, 6.490598792654666e-2, -1.249992184872738e-1)
_JP1 = (-4.878788132172128e-9, 6.009061827883699e-7, -4.541343896997497e-5, 1.937383947804541e-3, -3.405537384615824e-2)
_MO1 = (6.913942741265801e-2, -2.284801500053359e-1, 3.138238455499697e-1, -2.102302420403875e-1,
        5.435364690523026e-3, 1.493389585089498e-1, 4.976029650847191e-6, 7.978845453073848e-1)
_PH1 = (-4.497014141919556e1, 5.073465654089319e1, -2.485774108720340e1, 7.222973196770240e0,
        -1.544842782180211e0, 3.503787691653334e-1, -1.637986776941202e-1, 3.749989509080821e-1)


def _j0f(x):
    y = torch.abs(x)
    z = y * y
    tiny = 1.0 - 0.25 * z
    small = (z - 5.78318596294678452118) * _polevl(z, _JP)
    ys = torch.clamp(y, min=1e-30)
    q = 1.0 / ys
    p = torch.sqrt(q) * _polevl(q, _MO)
    yn = q * _polevl(q * q, _PH) - 0.7853981633974483096
    big = p * torch.cos(yn + y)
    return torch.where(y <= 2.0, torch.where(y < 1.0e-3, tiny, small), big)


def _j1f(x):
    y = torch.abs(x)
    z = y * y
    small = (z - 14.6819706421238932572) * y * _polevl(z, _JP1)
    ys = torch.clamp(y, min=1e-30)
    q = 1.0 / ys
    p = torch.sqrt(q) * _polevl(q, _MO1)
    yn = q * _polevl(q * q, _PH1) - 2.35619449019234492885
    big = p * torch.cos(yn + y)
    return torch.sign(x) * torch.where(y <= 2.0, small, big)


class _BesselJ0(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return _j0f(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return -_j1f(x) * dy            # tensorflow/python/ops/math_grad.py: _BesselJ0Grad


def _log(x):
    return torch.log(_t(x))


math = types.SimpleNamespace(
    exp=exp, sqrt=sqrt, log=_log, tanh=tanh, square=square, abs=abs, pow=pow,
    atanh=lambda x: torch.atanh(_t(x)), logical_and=logical_and,
    is_finite=lambda x: torch.isfinite(_t(x)),
    reduce_std=lambda x, axis=None: torch.std(_t(x), dim=axis, unbiased=False),
    special=types.SimpleNamespace(bessel_j0=lambda x: _BesselJ0.apply(_t(x)),
                                  bessel_j1=lambda x: _j1f(_t(x))),
)

nn = types.SimpleNamespace(sigmoid=lambda x: torch.sigmoid(_t(x)), relu=lambda x: torch.relu(_t(x)))


# ---------------------------------------------------------------- random (recorded)
class _Random:
    def __init__(self):
        self.LOG = []

    def set_seed(self, seed):
        torch.manual_seed(seed)

    def normal(self, shape, mean=0.0, stddev=1.0, dtype=float32):
        z = torch.randn(_shape_arg(shape), dtype=dtype)
        self.LOG.append(('normal', z.clone()))
        return z * stddev + mean

    def uniform(self, shape, minval=0, maxval=None, dtype=float32):
        shp = _shape_arg(shape)
        if dtype in (int32, int64):
            r = torch.randint(int(minval), int(maxval), shp, dtype=dtype)
            self.LOG.append(('uniform_int', r.clone()))
            return r
        maxval = 1.0 if maxval is None else maxval
        u = torch.rand(shp, dtype=dtype)
        self.LOG.append(('uniform', u.clone()))
        return u * (maxval - minval) + minval

    def shuffle(self, x):
        perm = torch.randperm(x.shape[0])
        self.LOG.append(('perm', perm.clone()))
        return x[perm]


random = _Random()


# ---------------------------------------------------------------- autodiff
class GradientTape:
    def __init__(self, persistent=False):
        self._persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, x):
        x.requires_grad_(True)

    def gradient(self, target, sources):
        single = isinstance(sources, torch.Tensor)
        srcs = [sources] if single else list(sources)
        g = torch.autograd.grad(target.sum(), srcs, retain_graph=self._persistent, allow_unused=True)
        return g[0] if single else list(g)


# ---------------------------------------------------------------- keras (only what the hot path needs)
keras = types.ModuleType('tensorflow.keras')
keras.layers = types.ModuleType('tensorflow.keras.layers')


class _Layer:
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


keras.layers.Layer = _Layer


# ---- a minimal functional API: what EncoderTrainer.create_encoder (model.py:122-223) touches.  Layers applied to the
# symbolic tensors of keras.layers.Input build a graph; keras.Model(inputs, outputs) replays it on real tensors.
class _Sym:
    """Symbolic tensor: `layer` applied to `parents` (output `index` of it), or a placeholder (layer None)."""

    def __init__(self, layer=None, parents=(), index=None, last_dim=None):
        self.layer, self.parents, self.index = layer, tuple(parents), index
        self.shape = (None, None, None, None, last_dim)


def _is_sym(x):
    return isinstance(x, _Sym) or (isinstance(x, (list, tuple)) and len(x) > 0 and all(isinstance(v, _Sym) for v in x))


class _GraphLayer:
    """Base of the functional-API layers: symbolic inputs defer, real inputs compute."""
    created = []                                       # creation order, so a test can install known weights

    def __call__(self, inputs):
        if _is_sym(inputs):
            parents = inputs if isinstance(inputs, (list, tuple)) else [inputs]
            return _Sym(self, parents, None, self._out_dim(parents))
        return self.compute(inputs)

    def _out_dim(self, parents):
        return parents[0].shape[-1]


class _Conv3D(_GraphLayer):
    """keras.layers.Conv3D on channels-last [B,X,Y,Z,C]: kernel [kx,ky,kz,C_in,C_out], padding 'valid' | 'same'
    (odd kernels), optional activation ('relu', None or a callable)."""

    def __init__(self, filters, kernel_size, padding='valid', kernel_initializer=None, bias_initializer=None,
                 activation=None):
        self.filters, self.kernel_size, self.padding = int(filters), tuple(kernel_size), padding
        self.kernel_initializer, self.bias_initializer, self.activation = kernel_initializer, bias_initializer, activation
        self.kernel = self.bias = None
        _GraphLayer.created.append(self)

    def _out_dim(self, parents):
        return self.filters

    def build(self, c_in):
        shape = self.kernel_size + (c_in, self.filters)
        init = self.kernel_initializer or (lambda shp: torch.zeros(shp))
        self.kernel = init(shape).float()
        self.bias = (self.bias_initializer((self.filters,)) if self.bias_initializer else torch.zeros(self.filters)).float()

    def compute(self, x):
        x = _t(x)
        if self.kernel is None:
            self.build(x.shape[-1])
        w = self.kernel.permute(4, 3, 0, 1, 2)                       # [C_out, C_in, kx, ky, kz]
        pad = tuple(k // 2 for k in self.kernel_size) if self.padding == 'same' else 0
        y = torch.nn.functional.conv3d(x.permute(0, 4, 1, 2, 3), w, self.bias, padding=pad).permute(0, 2, 3, 4, 1)
        act = self.activation
        if act is None or act == 'linear':
            return y
        if callable(act):
            return act(y)
        return {'relu': torch.relu, 'gelu': torch.nn.functional.gelu, 'tanh': torch.tanh}[act](y)


class _Activation(_GraphLayer):
    def __init__(self, activation):
        self.activation = activation

    def compute(self, x):
        return {'relu': torch.relu, 'gelu': torch.nn.functional.gelu, 'tanh': torch.tanh}[self.activation](_t(x))


class _Lambda(_GraphLayer):
    def __init__(self, fn):
        self.fn = fn

    def compute(self, x):
        return self.fn(x)


def _Input(shape=None, ragged=False, **kw):
    return _Sym(None, (), None, shape[-1] if shape else None)


class _Model(_GraphLayer):
    """keras.Model(inputs, outputs): callable on real tensors (list or single), or on symbolic tensors (nesting)."""

    def __init__(self, inputs=None, outputs=None):
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.single_output = not isinstance(outputs, (list, tuple))
        self.outputs = [outputs] if self.single_output else list(outputs)

    def __call__(self, inputs):
        if _is_sym(inputs):
            parents = inputs if isinstance(inputs, (list, tuple)) else [inputs]
            outs = [_Sym(self, parents, i, o.shape[-1]) for i, o in enumerate(self.outputs)]
            return outs[0] if self.single_output else outs
        return self.compute(inputs)

    def compute(self, inputs):
        vals = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        memo = {id(s): _t(v) for s, v in zip(self.inputs, vals)}
        nested = {}

        def ev(sym):
            if id(sym) in memo:
                return memo[id(sym)]
            if sym.layer is None:
                raise ValueError('unbound keras Input')
            args = [ev(p) for p in sym.parents]
            if isinstance(sym.layer, _Model):
                key = (id(sym.layer), tuple(id(p) for p in sym.parents))
                if key not in nested:
                    res = sym.layer.compute(args)
                    nested[key] = res if isinstance(res, (list, tuple)) else [res]
                out = nested[key][sym.index]
            else:
                out = sym.layer.compute(args[0] if len(args) == 1 else args)
            memo[id(sym)] = out
            return out

        outs = [ev(o) for o in self.outputs]
        return outs[0] if self.single_output else outs

    predict = compute


keras.layers.Conv3D = _Conv3D
keras.layers.Activation = _Activation
keras.layers.Lambda = _Lambda
keras.layers.Input = _Input
keras.layers.created = _GraphLayer.created
keras.Model = _Model
keras.backend = types.SimpleNamespace(print_tensor=lambda x, *a, **k: print(x))


def _he_normal():
    def init(shape):                                   # fan_in = receptive field x C_in; truncated normal, as keras
        fan_in = 1
        for d in shape[:-1]:
            fan_in *= d
        std = (2.0 / fan_in) ** 0.5 / 0.87962566103423978
        return torch.nn.init.trunc_normal_(torch.empty(shape), 0.0, std, -2 * std, 2 * std)
    return init


keras.initializers = types.SimpleNamespace(
    HeNormal=_he_normal,
    RandomNormal=lambda stddev=0.05, mean=0.0: (lambda shape: torch.randn(shape) * stddev + mean),
    Constant=lambda value: (lambda shape: torch.full(shape, float(value))),
    constant=lambda value: (lambda shape: torch.as_tensor(np.asarray(value, dtype=np.float32)).reshape(shape)))
sys.modules['tensorflow.keras'] = keras
sys.modules['tensorflow.keras.layers'] = keras.layers
