"""Parameter storage for the module mirrors.

Each module keeps ONE flat fp32 buffer for its parameters (and one for gradients); the named tensors of the
reference's parameter tree are views into it.  The flat layout is what makes the optimizer one
kernel launch per module (``arcvae_adam_step``) and the data-parallel gradient exchange one all-reduce per bucket.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

Spec = List[Tuple[str, Tuple[int, ...], str]]  # (dotted name, shape, init kind)


class ParamGroup:
    """``module.fc_out.weight`` style access, like an MLX sub-module."""

    def __init__(self, tensors: Dict[str, torch.Tensor]):
        self.__dict__.update(tensors)
        self._names = list(tensors)

    def items(self):
        return [(n, getattr(self, n)) for n in self._names]


def _align(n, a=64):
    return (n + a - 1) // a * a


AUX = 64  # floats appended to every flat buffer (not parameters: see FlatParams.aux)


class FlatParams:
    """Flat buffer + named views.  Every tensor starts at a 256-byte boundary.

    ``storage`` = ``flat`` (the parameters, what Adam and the clip see) followed by ``aux``, AUX spare floats.  The
    gradient copy of the decoder uses its aux tail to carry this rank's partial reconstruction loss through the SAME
    all-reduce as the gradients (data parallelism: no extra collective for the loss values)."""

    def __init__(self, spec: Spec, device):
        self.spec = spec
        self.offsets = OrderedDict()
        off = 0
        for name, shape, _ in spec:
            self.offsets[name] = (off, shape)
            off += _align(int(math.prod(shape)))
        self.numel = off
        self._alloc(torch.zeros(off + AUX, dtype=torch.float32, device=device))

    def _alloc(self, storage):
        self.storage = storage
        self.flat = storage[:self.numel]
        self.aux = storage[self.numel:]
        self.views = OrderedDict((n, self.flat[o:o + math.prod(s)].view(*s)) for n, (o, s) in self.offsets.items())

    def like(self):
        other = FlatParams.__new__(FlatParams)
        other.spec, other.offsets, other.numel = self.spec, self.offsets, self.numel
        other._alloc(torch.zeros_like(self.storage))
        return other

    def tree(self):
        out: dict = {}
        for name, v in self.views.items():
            mod, leaf = name.rsplit(".", 1)
            out.setdefault(mod, {})[leaf] = v
        return out


def init_mlx_style(fp: FlatParams, generator: torch.Generator):
    """MLX default initialisers: Linear U(-1/sqrt(in), 1/sqrt(in)) for weight and bias;
    LSTM U(-1/sqrt(H), 1/sqrt(H)); Embedding N(0, 1/E); encoder fc_logvar.bias = 0.35 (models/encoder.py:71-74)."""
    cpu = {}
    for name, shape, kind in fp.spec:
        if kind.startswith("uniform:"):
            k = 1.0 / math.sqrt(float(kind.split(":")[1]))
            t = (torch.rand(shape, generator=generator, dtype=torch.float64) * 2 - 1) * k
        elif kind.startswith("normal:"):
            t = torch.randn(shape, generator=generator, dtype=torch.float64) * math.sqrt(1.0 / float(kind.split(":")[1]))
        elif kind.startswith("const:"):
            t = torch.full(shape, float(kind.split(":")[1]), dtype=torch.float64)
        else:
            raise ValueError(kind)
        cpu[name] = t.to(torch.float32)
    for name, t in cpu.items():
        fp.views[name].copy_(t)


def encoder_spec(V, E, H, L, C, NL) -> Spec:
    """Order = construction order in models/encoder.py:46-74 (so a shared seeded generator reproduces the oracle)."""
    s: Spec = [("embedding.weight", (V, E), f"normal:{E}")]
    for i in range(NL):
        D = E if i == 0 else H
        s += [(f"lstm_layer_{i}.Wx", (4 * H, D), f"uniform:{H}"), (f"lstm_layer_{i}.Wh", (4 * H, H), f"uniform:{H}"),
              (f"lstm_layer_{i}.bias", (4 * H,), f"uniform:{H}")]
    s += [("condition_fc.weight", (H, C), f"uniform:{C}"), ("condition_fc.bias", (H,), f"uniform:{C}"),
          ("fc_mu.weight", (L, 2 * H), f"uniform:{2 * H}"), ("fc_mu.bias", (L,), f"uniform:{2 * H}"),
          ("fc_logvar_hidden.weight", (2 * H, 2 * H), f"uniform:{2 * H}"),
          ("fc_logvar_hidden.bias", (2 * H,), f"uniform:{2 * H}"),
          ("fc_logvar.weight", (L, 2 * H), f"uniform:{2 * H}"), ("fc_logvar.bias", (L,), "const:0.35")]
    return s


def decoder_spec(V, E, H, L, C, NL) -> Spec:
    """Order = construction order in models/decoder.py:51-73."""
    s: Spec = [("z_to_hidden.weight", (H, L), f"uniform:{L}"), ("z_to_hidden.bias", (H,), f"uniform:{L}"),
               ("condition_to_hidden.weight", (H, C), f"uniform:{C}"), ("condition_to_hidden.bias", (H,), f"uniform:{C}"),
               ("embedding.weight", (V, E), f"normal:{E}")]
    for i in range(NL):
        D = E + C if i == 0 else H
        s += [(f"lstm_layer_{i}.Wx", (4 * H, D), f"uniform:{H}"), (f"lstm_layer_{i}.Wh", (4 * H, H), f"uniform:{H}"),
              (f"lstm_layer_{i}.bias", (4 * H,), f"uniform:{H}")]
    s += [("fc_out.weight", (V, H), f"uniform:{H}"), ("fc_out.bias", (V,), f"uniform:{H}")]
    return s


class Workspace:
    """Grow-only cache of device byte buffers keyed by name (tapes, scratch)."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}

    def get(self, name: str, nbytes: int) -> torch.Tensor:
        b = self.bufs.get(name)
        if b is None or b.numel() < nbytes:
            self.bufs.pop(name, None)
            b = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self.bufs[name] = b
        return b
