"""``mlx.nn`` stand-in (see mlx/__init__.py): Module, Linear, Embedding, LSTM with MLX's parameter names, shapes,
initialisers and call semantics."""
from __future__ import annotations

import math

import torch

from .. import core as mx


class Module(dict):
    """MLX ``nn.Module`` is a ``dict`` subclass: arrays and sub-modules (and containers of them) assigned as attributes
    live IN the dict; everything else is an ordinary attribute."""

    def __init__(self):
        dict.__init__(self)
        object.__setattr__(self, "_no_grad", set())
        object.__setattr__(self, "_training", True)

    def __getattr__(self, key):
        if key in self:
            return self[key]
        raise AttributeError(f"{type(self).__name__!r} has no attribute {key!r}")

    def __setattr__(self, key, val):
        if isinstance(val, (mx.array, dict, list, tuple)):
            self.__dict__.pop(key, None)
            self[key] = val
        else:
            if key in self:
                del self[key]
            object.__setattr__(self, key, val)

    def __call__(self, *a, **k):
        raise NotImplementedError

    @staticmethod
    def _params_of(tree):
        out = {}
        for k, v in tree.items():
            if isinstance(v, mx.array):
                out[k] = v
            elif isinstance(v, dict):
                sub = Module._params_of(v)
                out[k] = sub
        return out

    def parameters(self):
        return Module._params_of(self)

    def trainable_parameters(self):
        return self.parameters()

    def children(self):
        return {k: v for k, v in self.items() if isinstance(v, Module)}

    def update(self, parameters):
        def apply(dst, src):
            for k, v in src.items():
                if isinstance(v, dict):
                    apply(dst[k], v)
                else:
                    dict.__setitem__(dst, k, v)
        apply(self, parameters)
        return self

    def train(self, mode=True):
        object.__setattr__(self, "_training", mode)
        return self

    def eval(self):
        return self.train(False)


def _uniform(shape, bound):
    return mx.random.uniform(-bound, bound, shape)


class Linear(Module):
    """weight [out, in], bias [out], both U(-1/sqrt(in), 1/sqrt(in)); y = addmm(bias, x, weight.T)."""

    def __init__(self, input_dims: int, output_dims: int, bias: bool = True):
        super().__init__()
        scale = math.sqrt(1.0 / input_dims)
        self.weight = _uniform((output_dims, input_dims), scale)
        if bias:
            self.bias = _uniform((output_dims,), scale)

    def __call__(self, x):
        if "bias" in self:
            return mx.addmm(self["bias"], x, self["weight"].T)
        return x @ self["weight"].T


class Embedding(Module):
    """weight [num_embeddings, dims] ~ N(0, 1/dims); y = weight[x]."""

    def __init__(self, num_embeddings: int, dims: int):
        super().__init__()
        self.weight = mx.random.normal((num_embeddings, dims), scale=math.sqrt(1.0 / dims))

    def __call__(self, x):
        return self["weight"][mx.array(x).long() if not isinstance(x, torch.Tensor) else x.long()]


class LSTM(Module):
    """Single-layer LSTM.  Wx [4H, D], Wh [4H, H], bias [4H] (ONE bias), U(-1/sqrt(H), 1/sqrt(H)).

    ``__call__(x, hidden=None, cell=None)``: the input projection of all positions at once (bias included), then a
    Python loop over axis -2; the recurrent product is added only when ``hidden`` is not None (so the first step of a
    fresh call has no Wh term); gate order i, f, g, o; ``cell = f*cell + i*g`` once a cell exists, else ``i*g``;
    returns (all hidden states, all cell states), each [..., T, H]."""

    def __init__(self, input_size: int, hidden_size: int, bias: bool = True):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        scale = 1.0 / math.sqrt(hidden_size)
        self.Wx = _uniform((4 * hidden_size, input_size), scale)
        self.Wh = _uniform((4 * hidden_size, hidden_size), scale)
        self.bias = _uniform((4 * hidden_size,), scale) if bias else None

    def __call__(self, x, hidden=None, cell=None):
        bias = self["bias"] if "bias" in self else None
        x = mx.addmm(bias, x, self["Wx"].T) if bias is not None else x @ self["Wx"].T
        hs, cs = [], []
        for idx in range(x.shape[-2]):
            ifgo = x[..., idx, :]
            if hidden is not None:
                ifgo = ifgo + hidden @ self["Wh"].T
            i, f, g, o = mx.split(ifgo, 4, axis=-1)
            i, f, g, o = mx.sigmoid(i), mx.sigmoid(f), mx.tanh(g), mx.sigmoid(o)
            cell = f * cell + i * g if cell is not None else i * g
            hidden = o * mx.tanh(cell)
            cs.append(cell)
            hs.append(hidden)
        return mx.stack(hs, axis=-2), mx.stack(cs, axis=-2)


def value_and_grad(model, fn):
    """``nn.value_and_grad(model, fn)``: gradients w.r.t. model.trainable_parameters()."""
    def inner(*args, **kwargs):
        return mx.value_and_grad(lambda m, *a, **k: fn(*a, **k), argnums=0)(model, *args, **kwargs)
    return inner
