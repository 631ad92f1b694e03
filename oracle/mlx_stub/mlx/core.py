"""``mlx.core`` stand-in (see mlx/__init__.py).  Arrays are torch CPU tensors of class ``array``.

Stub-only switches (never used by the reference code itself):
  set_default_float(torch.float64)  run the reference in fp64 so it can be compared with the fp64 oracle at 1e-12
  random.inject(seq)                queue of tensors returned by the next ``mx.random.normal`` calls (the reference never
                                    seeds mx.random, so its draws must be injected to be comparable)
"""
from __future__ import annotations

import builtins
from typing import Sequence

import numpy as np
import torch

_FLOAT = torch.float32


def set_default_float(dt):
    global _FLOAT
    _FLOAT = dt


class Dtype:
    def __init__(self, name, kind):
        self.name, self.kind = name, kind

    @property
    def torch(self):
        return {"f": _FLOAT, "i": torch.int64, "b": torch.bool}[self.kind]

    def __repr__(self):
        return f"mlx.core.{self.name}"


float32, float16, bfloat16 = Dtype("float32", "f"), Dtype("float16", "f"), Dtype("bfloat16", "f")
uint32, int32, int64, uint8 = Dtype("uint32", "i"), Dtype("int32", "i"), Dtype("int64", "i"), Dtype("uint8", "i")
bool_ = Dtype("bool_", "b")


def _tdt(dt):
    if dt is None:
        return None
    return dt.torch if isinstance(dt, Dtype) else dt


class array(torch.Tensor):
    """``mx.array``.  Floats live in the default float dtype, every integer type in int64 (torch indexes with it)."""

    @staticmethod
    def __new__(cls, data=0.0, dtype=None):
        if isinstance(data, torch.Tensor):
            t = data.detach() if not data.requires_grad else data
        else:
            t = torch.as_tensor(np.asarray(data))
        want = _tdt(dtype)
        if want is None:
            want = _FLOAT if t.is_floating_point() else (torch.bool if t.dtype == torch.bool else torch.int64)
        if t.dtype != want:
            t = t.to(want)
        return t.as_subclass(cls)

    def astype(self, dt):
        return self.to(_tdt(dt))

    @property
    def T(self):                                   # mx.array.T: full transpose (2-D here)
        return self.transpose(-1, -2) if self.ndim >= 2 else self


def _a(x):
    return x if isinstance(x, array) else array(x)


def _axis_kw(axis, keepdims):
    kw = {}
    if axis is not None:
        kw["dim"] = axis
        kw["keepdim"] = keepdims
    return kw


# ---- creation / shape ----------------------------------------------------------------------------------------------
def zeros(shape, dtype=float32):
    return torch.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=_tdt(dtype)).as_subclass(array)


def ones(shape, dtype=float32):
    return torch.ones(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=_tdt(dtype)).as_subclass(array)


def zeros_like(a):
    return torch.zeros_like(_a(a)).as_subclass(array)


def ones_like(a):
    return torch.ones_like(_a(a)).as_subclass(array)


def reshape(a, shape):
    return _a(a).reshape(tuple(shape))


def expand_dims(a, axis):
    return _a(a).unsqueeze(axis)


def squeeze(a, axis=None):
    return _a(a).squeeze() if axis is None else _a(a).squeeze(axis)


def concatenate(arrays: Sequence, axis=0):
    return torch.cat([_a(x) for x in arrays], dim=axis)


def stack(arrays: Sequence, axis=0):
    return torch.stack([_a(x) for x in arrays], dim=axis)


def split(a, parts, axis=0):
    a = _a(a)
    return list(torch.split(a, a.shape[axis] // parts, dim=axis))


def repeat(a, repeats, axis=None):
    return torch.repeat_interleave(_a(a), repeats, dim=axis)


def take_along_axis(a, indices, axis):
    return torch.take_along_dim(_a(a), _a(indices).long(), dim=axis)


def transpose(a, axes=None):
    a = _a(a)
    return a.permute(*reversed(range(a.ndim))) if axes is None else a.permute(*axes)


# ---- element-wise -----------------------------------------------------------------------------------------------------
def exp(a): return torch.exp(_a(a))
def log(a): return torch.log(_a(a))
def tanh(a): return torch.tanh(_a(a))
def sigmoid(a): return torch.sigmoid(_a(a))
def sqrt(a): return torch.sqrt(_a(a))
def square(a): return _a(a) * _a(a)
def abs(a): return torch.abs(_a(a))           # noqa: A001 (mirrors mx.abs)
def logical_or(a, b): return torch.logical_or(_a(a), _a(b))
def logical_and(a, b): return torch.logical_and(_a(a), _a(b))
def where(c, a, b): return torch.where(_a(c), _a(a), _a(b))
def isnan(a): return torch.isnan(_a(a))
def isinf(a): return torch.isinf(_a(a))
def stop_gradient(a): return _a(a).detach()


class _Maximum(torch.autograd.Function):
    """MLX ``Maximum::vjp``: the cotangent goes to ``a`` where ``a > b`` and to ``b`` elsewhere (ties -> b)."""

    @staticmethod
    def forward(ctx, a, b):
        mask = a > b
        ctx.save_for_backward(mask)
        ctx.sa, ctx.sb = a.shape, b.shape
        return torch.where(mask, a, b)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return _unbroadcast(g * mask, ctx.sa), _unbroadcast(g * (~mask), ctx.sb)


class _Minimum(torch.autograd.Function):
    """MLX ``Minimum::vjp``: cotangent to ``a`` where ``a < b``, else to ``b``."""

    @staticmethod
    def forward(ctx, a, b):
        mask = a < b
        ctx.save_for_backward(mask)
        ctx.sa, ctx.sb = a.shape, b.shape
        return torch.where(mask, a, b)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return _unbroadcast(g * mask, ctx.sa), _unbroadcast(g * (~mask), ctx.sb)


def _unbroadcast(g, shape):
    while g.ndim > len(shape):
        g = g.sum(0)
    for i, s in enumerate(shape):
        if s == 1 and g.shape[i] != 1:
            g = g.sum(i, keepdim=True)
    return g


def _pair(a, b):
    a, b = _a(a), _a(b)
    if a.dtype != b.dtype:
        dt = torch.promote_types(a.dtype, b.dtype)
        a, b = a.to(dt), b.to(dt)
    return a, b


def maximum(a, b):
    a, b = _pair(a, b)
    return _Maximum.apply(a, b).as_subclass(array)


def minimum(a, b):
    a, b = _pair(a, b)
    return _Minimum.apply(a, b).as_subclass(array)


def clip(a, a_min, a_max):
    """mx.clip = minimum(maximum(a, a_min), a_max) (either bound may be None)."""
    out = _a(a)
    if a_min is not None:
        out = maximum(out, a_min)
    if a_max is not None:
        out = minimum(out, a_max)
    return out


# ---- reductions -------------------------------------------------------------------------------------------------------
def sum(a, axis=None, keepdims=False):        # noqa: A001
    return torch.sum(_a(a), **_axis_kw(axis, keepdims))


def mean(a, axis=None, keepdims=False):
    return torch.mean(_a(a), **_axis_kw(axis, keepdims))


def max(a, axis=None, keepdims=False):        # noqa: A001
    a = _a(a)
    return torch.amax(a, **_axis_kw(axis, keepdims)) if axis is not None else a.max()


def min(a, axis=None, keepdims=False):        # noqa: A001
    a = _a(a)
    return torch.amin(a, **_axis_kw(axis, keepdims)) if axis is not None else a.min()


def all(a, axis=None, keepdims=False):        # noqa: A001
    return torch.all(_a(a), **_axis_kw(axis, keepdims))


def any(a, axis=None, keepdims=False):        # noqa: A001
    return torch.any(_a(a), **_axis_kw(axis, keepdims))


def argmax(a, axis=None, keepdims=False):
    """First (lowest) index on ties, as mx.argmax; result uint32 in MLX, int64 here."""
    a = _a(a)
    if axis is None:
        return torch.argmax(a.reshape(-1)).as_subclass(array)
    return torch.argmax(a, dim=axis, keepdim=keepdims)


def softmax(a, axis=-1):
    return torch.softmax(_a(a), dim=axis)


def logsumexp(a, axis=None, keepdims=False):
    return torch.logsumexp(_a(a), dim=axis, keepdim=keepdims)


def addmm(c, a, b, alpha=1.0, beta=1.0):
    return alpha * (_a(a) @ _a(b)) + beta * _a(c)


def matmul(a, b):
    return _a(a) @ _a(b)


def eval(*args):                               # noqa: A001 — MLX is lazy; torch is eager: nothing to do
    return None


# ---- random -----------------------------------------------------------------------------------------------------------
class _Random:
    def __init__(self):
        self._queue = []
        self._gen = torch.Generator().manual_seed(0)

    def inject(self, tensors):
        """Stub-only: the next ``normal`` calls return these tensors (shape-checked) instead of fresh draws."""
        self._queue.extend(tensors)

    def pending(self):
        return len(self._queue)

    def seed(self, s):
        self._gen.manual_seed(int(s))

    def key(self, s):
        return int(s)

    def normal(self, shape=(), dtype=float32, loc=0.0, scale=1.0, key=None):
        shape = tuple(shape)
        if self._queue:
            t = _a(self._queue.pop(0)).to(_tdt(dtype))
            assert tuple(t.shape) == shape, (tuple(t.shape), shape)
            return (t * scale + loc).as_subclass(array)
        return (torch.randn(shape, generator=self._gen, dtype=torch.float64) * scale + loc).to(_tdt(dtype)).as_subclass(array)

    def uniform(self, low=0.0, high=1.0, shape=(), dtype=float32, key=None):
        u = torch.rand(tuple(shape), generator=self._gen, dtype=torch.float64)
        return (u * (high - low) + low).to(_tdt(dtype)).as_subclass(array)

    def randint(self, low, high, shape=(), dtype=int32, key=None):
        return torch.randint(int(low), int(high), tuple(shape), generator=self._gen).as_subclass(array)


random = _Random()


# ---- transforms ---------------------------------------------------------------------------------------------------------
def _leaves_inplace(tree, fn):
    """Walk dict / list containers IN PLACE, replacing every array leaf by fn(leaf); returns the new leaves in order.
    (MLX's value_and_grad fills the tracers into the caller's own containers — which is why an nn.Module, a dict
    subclass, keeps its methods inside the differentiated function.)"""
    out = []
    if isinstance(tree, dict):
        for k in list(tree.keys()):
            v = tree[k]
            if isinstance(v, array):
                nv = fn(v)
                dict.__setitem__(tree, k, nv)
                out.append(nv)
            elif isinstance(v, (dict, list)):
                out.extend(_leaves_inplace(v, fn))
    elif isinstance(tree, list):
        for i, v in enumerate(tree):
            if isinstance(v, array):
                tree[i] = fn(v)
                out.append(tree[i])
            elif isinstance(v, (dict, list)):
                out.extend(_leaves_inplace(v, fn))
    return out


def _like_tree(tree, it):
    """Plain nested dict / list with the structure of ``tree`` and leaves drawn from ``it`` (gradients are returned as
    plain containers, never as nn.Module objects)."""
    if isinstance(tree, dict):
        out = {}
        for k, v in tree.items():
            if isinstance(v, array):
                out[k] = next(it)
            elif isinstance(v, (dict, list)):
                out[k] = _like_tree(v, it)
        return out
    if isinstance(tree, list):
        return [next(it) if isinstance(v, array) else _like_tree(v, it) for v in tree if isinstance(v, (array, dict, list))]
    raise TypeError(type(tree))


def value_and_grad(fun, argnums=0, argnames=None):
    single = isinstance(argnums, int)
    nums = [argnums] if single else list(argnums)

    def wrapped(*args, **kwargs):
        args = list(args)
        groups = []
        for n in nums:
            if isinstance(args[n], array):
                args[n] = args[n].detach().clone().requires_grad_(True).as_subclass(array)
                groups.append([args[n]])
            else:
                groups.append(_leaves_inplace(args[n], lambda v: v.detach().clone().requires_grad_(True).as_subclass(array)))
        value = fun(*args, **kwargs)
        first = value[0] if isinstance(value, (tuple, list)) else value
        flat = [leaf for g in groups for leaf in g]
        gs = torch.autograd.grad(first, flat, allow_unused=True)
        gs = [(g if g is not None else torch.zeros_like(p)).detach().as_subclass(array) for g, p in zip(gs, flat)]
        grads, pos = [], 0
        for n, g in zip(nums, groups):
            chunk = gs[pos:pos + len(g)]
            pos += len(g)
            grads.append(chunk[0] if isinstance(args[n], array) else _like_tree(args[n], iter(chunk)))
        for n in nums:                       # leave the caller's containers holding plain (non-tracked) leaves
            if not isinstance(args[n], array):
                _leaves_inplace(args[n], lambda v: v.detach().as_subclass(array))
        value = value.detach() if isinstance(value, torch.Tensor) else value
        return value, (grads[0] if single else tuple(grads))

    return wrapped


def grad(fun, argnums=0, argnames=None):
    vg = value_and_grad(fun, argnums, argnames)
    return lambda *a, **k: vg(*a, **k)[1]
