"""Base class of the module mirrors: owns the flat parameter / gradient buffers and the ctypes views of them."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _lib
from ._params import FlatParams, ParamGroup, Workspace, init_mlx_style


def default_device():
    if not torch.cuda.is_available():
        raise _lib.ArcvaeError("mlx_vae_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def parse_precision(p) -> int:
    if p in (0, "fp32", "float32"):
        return _lib.PREC_FP32
    if p in (1, "bf16", "bfloat16"):
        return _lib.PREC_BF16
    raise ValueError(f"precision must be 'fp32' or 'bf16', got {p!r}")


class Module:
    _struct = None  # ctypes struct class

    def _setup(self, spec, dims: Dict[str, int], device, seed: Optional[int], precision):
        self.device = torch.device(device) if device is not None else default_device()
        self.precision = parse_precision(precision)
        self._dims = _lib.Dims(**dims)
        self.params = FlatParams(spec, self.device)
        gen = torch.Generator()
        # seed=None: fresh entropy WITHOUT touching torch's global generator (torch.seed() would reseed it)
        gen.manual_seed(int(seed) if seed is not None else int.from_bytes(os.urandom(4), "little") % (2 ** 31))
        init_mlx_style(self.params, gen)
        self.grads = self.params.like()
        self.ws = Workspace(self.device)
        self._attach_groups()
        self._cparams = self._make_struct(self.params)
        self._cgrads = self._make_struct(self.grads)

    def _attach_groups(self):
        for mod, leaves in self.params.tree().items():
            setattr(self, mod, ParamGroup(leaves))

    def _make_struct(self, fp: FlatParams):
        s = self._struct()
        v = fp.views
        for name, t in v.items():
            mod, leaf = name.rsplit(".", 1)
            if mod.startswith("lstm_layer_"):
                i = int(mod.rsplit("_", 1)[1])
                getattr(s, leaf)[i] = t.data_ptr()
            elif mod == "embedding":
                s.embedding = t.data_ptr()
            else:
                setattr(s, f"{mod}_{'w' if leaf == 'weight' else 'b'}", t.data_ptr())
        return s

    # ---- the reference's nn.Module surface (MLX modules are dicts of arrays / sub-modules) -----------------
    def parameters(self):
        """Nested dict keyed like the reference's parameter tree (``encoder.parameters()`` in trainer.py:329)."""
        return self.params.tree()

    def gradients(self):
        """Nested dict with the same keys: what ``mx.value_and_grad`` returns per module (trainer.py:305)."""
        return self.grads.tree()

    def state_dict(self):
        return dict(self.params.views)

    def load_parameters(self, tree):
        """Accepts a nested dict {module: {leaf: array}} or a flat {'module.leaf': array}; values may be torch / numpy."""
        flat = {}
        for k, v in tree.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    flat[f"{k}.{kk}"] = vv
            else:
                flat[k] = v
        missing = set(self.params.views) - set(flat)
        if missing:
            raise KeyError(f"missing parameters: {sorted(missing)}")
        for name, dst in self.params.views.items():
            src = torch.as_tensor(flat[name])
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{name}: shape {tuple(src.shape)} != {tuple(dst.shape)}")
            dst.copy_(src.to(device=self.device, dtype=torch.float32))
        return self

    def zero_grad(self):
        _lib.check(_lib.load().arcvae_zero(self.grads.storage.data_ptr(), self.grads.storage.numel() * 4, _lib.stream_ptr()))

    def num_parameters(self) -> int:
        import math
        return sum(math.prod(s) for _, (o, s) in self.params.offsets.items())

    # ---- helpers ---------------------------------------------------------------------------------------
    def _tokens(self, x) -> torch.Tensor:
        _lib.require_cuda(x)
        if x.dtype != torch.int32 or not x.is_contiguous():
            x = x.to(torch.int32).contiguous()
        return x

    def _f32(self, t) -> torch.Tensor:
        _lib.require_cuda(t)
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        return t
