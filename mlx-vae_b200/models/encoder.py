"""MLXEncoder — drop-in for models/encoder.py of the reference (same constructor, call and static method)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import _lib
from .._module import Module
from .._params import encoder_spec


class MLXEncoder(Module):
    """Conditional encoder of the AR-CVAE (models/encoder.py:4-132).

    ``encoder(x, conditions) -> (mu, logvar)`` with x [B,T] integer tokens and conditions [B,C] float32, both CUDA
    tensors.  ``dropout`` is accepted and unused, as in the reference (encoder.py:24).  The forward keeps a tape so
    that ``backward(dmu, dlogvar)`` can run BPTT; gradients accumulate into ``self.grads``.
    """
    _struct = _lib.EncoderParams

    def __init__(self, vocab_size: int, embedding_dim: int = 256, hidden_dim: int = 512, latent_dim: int = 200,
                 num_conditions: int = 6, num_layers: int = 3, dropout: float = 0.2, *, device=None,
                 seed: Optional[int] = None, precision="fp32"):
        self.vocab_size, self.embedding_dim, self.hidden_dim = vocab_size, embedding_dim, hidden_dim
        self.latent_dim, self.num_conditions, self.num_layers = latent_dim, num_conditions, num_layers
        self._setup(encoder_spec(vocab_size, embedding_dim, hidden_dim, latent_dim, num_conditions, num_layers),
                    dict(V=vocab_size, E=embedding_dim, H=hidden_dim, L=latent_dim, C=num_conditions, NL=num_layers,
                         pad_token=0, end_token=2), device, seed, precision)
        self._ctx = None
        self._last_bt = None

    def __call__(self, x: torch.Tensor, conditions: torch.Tensor, *, out=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``out=(mu, logvar)``: write into caller-owned [B,L] float32 buffers (used when the call is enqueued on a side
        stream: the caller allocates on its own stream, so the caching allocator never crosses streams)."""
        lib = _lib.load()
        x = self._tokens(x)
        cond = self._f32(conditions)
        B, T = x.shape
        if cond.shape != (B, self.num_conditions):
            raise ValueError(f"conditions must be [{B},{self.num_conditions}], got {tuple(cond.shape)}")
        nbytes = lib.arcvae_encoder_tape_bytes(self._dims, B, T)
        tape = self.ws.get("tape", nbytes)
        if out is not None:
            mu, logvar = out
        else:
            mu = torch.empty((B, self.latent_dim), dtype=torch.float32, device=self.device)
            logvar = torch.empty_like(mu)
        _lib.check(lib.arcvae_encoder_forward(self._dims, self._cparams, x.data_ptr(), cond.data_ptr(), B, T,
                                              mu.data_ptr(), logvar.data_ptr(), tape.data_ptr(), tape.numel(),
                                              self.precision, _lib.stream_ptr()))
        self._ctx = (B, T, cond, tape)
        self._last_bt = (B, T)
        return mu, logvar

    def backward(self, dmu: torch.Tensor, dlogvar: torch.Tensor):
        """Reverse pass of the last ``__call__`` (the encoder half of ``mx.value_and_grad``, trainer.py:292)."""
        if self._ctx is None:
            raise _lib.ArcvaeError("encoder.backward() without a forward")
        lib = _lib.load()
        B, T, cond, tape = self._ctx
        dmu, dlogvar = self._f32(dmu), self._f32(dlogvar)
        sbytes = lib.arcvae_encoder_scratch_bytes(self._dims, B, T)
        scratch = self.ws.get("scratch", sbytes)
        _lib.check(lib.arcvae_encoder_backward(self._dims, self._cparams, cond.data_ptr(), B, T, dmu.data_ptr(),
                                               dlogvar.data_ptr(), tape.data_ptr(), tape.numel(), self._cgrads,
                                               scratch.data_ptr(), scratch.numel(), self.precision, _lib.stream_ptr()))
        self._ctx = None

    def check(self):
        """Raise if a persistent cluster kernel of the last forward/backward reported a barrier time-out (the kernels use
        bounded waits instead of hanging).  Synchronises the stream; for tests and debugging."""
        tape = self.ws.bufs.get("tape")
        if tape is None or self._last_bt is None:
            return
        B, T = self._last_bt
        _lib.check(_lib.load().arcvae_encoder_check(self._dims, B, T, tape.data_ptr(), tape.numel(), self.precision,
                                                    _lib.stream_ptr()))

    @staticmethod
    def reparameterize(mu: torch.Tensor, logvar: torch.Tensor, eps: Optional[torch.Tensor] = None, *,
                       seed: int = 0, offset: int = 0) -> torch.Tensor:
        """z = mu + eps * exp(0.5 * logvar) (encoder.py:134-155).  ``eps=None`` draws N(0,1) on the device with
        Philox4x32-10 keyed by (seed, offset); pass ``eps`` to inject the draw (parity tests)."""
        lib = _lib.load()
        _lib.require_cuda(mu, logvar, eps)
        mu = mu.contiguous().float()
        logvar = logvar.contiguous().float()
        if eps is not None:
            eps = eps.contiguous().float()
        z = torch.empty_like(mu)
        B, L = mu.shape
        _lib.check(lib.arcvae_reparameterize(mu.data_ptr(), logvar.data_ptr(), _lib.ptr(eps), B, L, seed, offset,
                                             z.data_ptr(), _lib.stream_ptr()))
        return z
