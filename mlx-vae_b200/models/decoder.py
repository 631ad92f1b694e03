"""MLXAutoregressiveDecoder — drop-in for models/decoder.py of the reference."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .._module import Module
from .._params import decoder_spec


class MLXAutoregressiveDecoder(Module):
    """Autoregressive decoder (models/decoder.py:7-190).

    ``decoder(z, conditions, target_seq=None, max_length=80, teacher_forcing_ratio=0.5) -> logits [B,T,V]``.
    As in the reference, the LSTM layers are called without state at every position, so ``z`` does not influence
    the logits; it is accepted for signature compatibility.

    Teacher forcing: the reference draws ONE host coin per position for the whole batch,
    ``np.random.rand() < teacher_forcing_ratio`` (decoder.py:180).  This mirror draws the same coins from the same
    global NumPy RNG in the same order unless ``tf_mask`` (bool [T]) is given.
    """
    _struct = _lib.DecoderParams

    def __init__(self, vocab_size: int, embedding_dim: int = 256, hidden_dim: int = 512, latent_dim: int = 200,
                 num_conditions: int = 6, num_layers: int = 3, pad_token: int = 0, end_token: int = 2, *,
                 device=None, seed: Optional[int] = None, precision="fp32", carry_state: bool = False):
        self.vocab_size, self.embedding_dim, self.hidden_dim = vocab_size, embedding_dim, hidden_dim
        self.latent_dim, self.num_conditions, self.num_layers = latent_dim, num_conditions, num_layers
        self.pad_token, self.end_token = pad_token, end_token
        self._setup(decoder_spec(vocab_size, embedding_dim, hidden_dim, latent_dim, num_conditions, num_layers),
                    dict(V=vocab_size, E=embedding_dim, H=hidden_dim, L=latent_dim, C=num_conditions, NL=num_layers,
                         pad_token=pad_token, end_token=end_token), device, seed, precision)
        self._ctx = None
        self.last_inputs = None   # [B,T] tokens fed at each position by the last call (diagnostic)
        # carry_state=True is an EXTENSION, not the reference: the state of initialize_hidden_state is
        # consumed and carried across positions, z reaches the logits and backward() returns d total / d z
        self.carry_state = bool(carry_state)
        self.last_dz = None

    def initialize_hidden_state(self, z: torch.Tensor, conditions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """decoder.py:76-111.  The reference computes this and never uses it (F1); provided for API parity.
        Not on the hot path: two tiny products through the library's fp32 GEMM."""
        lib = _lib.load()
        z, c = self._f32(z), self._f32(conditions)
        B, H = z.shape[0], self.hidden_dim
        out = torch.empty((B, H), dtype=torch.float32, device=self.device)
        tmp = torch.empty_like(out)
        v = self.params.views
        s = _lib.stream_ptr()
        _lib.check(lib.arcvae_gemm_f32(0, 1, B, H, self.latent_dim, z.data_ptr(), self.latent_dim,
                                       v["z_to_hidden.weight"].data_ptr(), self.latent_dim, out.data_ptr(), H,
                                       v["z_to_hidden.bias"].data_ptr(), 0, s))
        _lib.check(lib.arcvae_gemm_f32(0, 1, B, H, self.num_conditions, c.data_ptr(), self.num_conditions,
                                       v["condition_to_hidden.weight"].data_ptr(), self.num_conditions, tmp.data_ptr(),
                                       H, v["condition_to_hidden.bias"].data_ptr(), 0, s))
        hidden_init = (out + tmp) / 2.0
        hidden = hidden_init.unsqueeze(0).repeat(self.num_layers, 1, 1)
        return hidden, torch.zeros_like(hidden)

    def __call__(self, z: torch.Tensor, conditions: torch.Tensor, target_seq: Optional[torch.Tensor] = None,
                 max_length: int = 80, teacher_forcing_ratio: float = 0.5, *,
                 tf_mask: Optional[Sequence[bool]] = None) -> torch.Tensor:
        lib = _lib.load()
        cond = self._f32(conditions)
        B = cond.shape[0]
        if target_seq is not None:
            target = self._tokens(target_seq)
            T = target.shape[1]                                   # decoder.py:137-140
        else:
            target, T = None, int(max_length)
        if tf_mask is None:
            # decoder.py:180 — `target_seq is not None and np.random.rand() < ratio`: short-circuit => no draw without target
            mask = np.zeros(T, dtype=np.uint8)
            if target is not None:
                for t in range(T):
                    mask[t] = 1 if np.random.rand() < teacher_forcing_ratio else 0
        else:
            mask = np.ascontiguousarray(np.asarray(tf_mask).astype(np.uint8))
            if mask.shape != (T,):
                raise ValueError(f"tf_mask must have shape ({T},)")
        logits_tm = torch.empty((T, B, self.vocab_size), dtype=torch.float32, device=self.device)
        inputs_tm = torch.empty((T, B), dtype=torch.int32, device=self.device)
        if self.carry_state:
            if z is None:
                raise ValueError("carry_state=True: the decoder state is initialised from z, which must be given")
            zz = self._f32(z)
            if zz.shape != (B, self.latent_dim):
                raise ValueError(f"z must be [{B},{self.latent_dim}], got {tuple(zz.shape)}")
            tape = self.ws.get("tape_cs", lib.arcvae_decoder_cs_tape_bytes(self._dims, B, T))
            _lib.check(lib.arcvae_decoder_cs_forward(self._dims, self._cparams, zz.data_ptr(), cond.data_ptr(),
                                                     _lib.ptr(target), mask.ctypes.data, B, T, logits_tm.data_ptr(),
                                                     inputs_tm.data_ptr(), tape.data_ptr(), tape.numel(), self.precision,
                                                     _lib.stream_ptr()))
            self._ctx = (B, T, cond, tape, zz)
        else:
            tape = self.ws.get("tape", lib.arcvae_decoder_tape_bytes(self._dims, B, T))
            _lib.check(lib.arcvae_decoder_forward(self._dims, self._cparams, cond.data_ptr(), _lib.ptr(target),
                                                  mask.ctypes.data, B, T, logits_tm.data_ptr(), inputs_tm.data_ptr(),
                                                  tape.data_ptr(), tape.numel(), self.precision, _lib.stream_ptr()))
            self._ctx = (B, T, cond, tape)
        self.last_inputs = inputs_tm.transpose(0, 1)
        self.last_tf_mask = mask.astype(bool)
        return logits_tm.transpose(0, 1)   # [B,T,V] view of the time-major buffer

    def ce_supported(self, B: int) -> bool:
        return (not self.carry_state) and bool(_lib.load().arcvae_decoder_ce_supported(self._dims, int(B), self.precision))

    def forward_ce(self, conditions: torch.Tensor, target_seq: torch.Tensor, ce_sum: torch.Tensor, ce_scale: float,
                   teacher_forcing_ratio: float = 0.5, *, tf_mask: Optional[Sequence[bool]] = None) -> None:
        """Training-step variant of ``__call__`` for the fused bf16 path: fc_out runs with the cross-entropy in its GEMM
        epilogue, so the fp32 logits [B,T,V] are never written.  ``ce_sum`` (a float64 device scalar view, e.g.
        ``stats[2L+2:2L+3]``) accumulates the CE sum; d logits (scaled by ``ce_scale`` = 1 / global positions) stay in the
        tape as bf16 for ``backward(None)``.  Same coins / feedback semantics as ``__call__`` (decoder.py:180-185)."""
        lib = _lib.load()
        cond = self._f32(conditions)
        target = self._tokens(target_seq)
        B, T = target.shape
        if tf_mask is None:
            mask = np.zeros(T, dtype=np.uint8)
            for t in range(T):
                mask[t] = 1 if np.random.rand() < teacher_forcing_ratio else 0
        else:
            mask = np.ascontiguousarray(np.asarray(tf_mask).astype(np.uint8))
            if mask.shape != (T,):
                raise ValueError(f"tf_mask must have shape ({T},)")
        tape = self.ws.get("tape", lib.arcvae_decoder_tape_bytes(self._dims, B, T))
        inputs_tm = torch.empty((T, B), dtype=torch.int32, device=self.device)
        _lib.check(lib.arcvae_decoder_forward_ce(self._dims, self._cparams, cond.data_ptr(), target.data_ptr(),
                                                 mask.ctypes.data, B, T, float(ce_scale), ce_sum.data_ptr(),
                                                 inputs_tm.data_ptr(), tape.data_ptr(), tape.numel(), self.precision,
                                                 _lib.stream_ptr()))
        self._ctx = (B, T, cond, tape)
        self.last_inputs = inputs_tm.transpose(0, 1)
        self.last_tf_mask = mask.astype(bool)

    def backward(self, dlogits: Optional[torch.Tensor]):
        """Reverse pass of the last call.  ``dlogits`` is [B,T,V]; the transposed view of a time-major [T,B,V]
        buffer (what the fused loss writes) is consumed in place, anything else is copied.  Returns None in the
        reference mode (z does not reach the logits) and d total / d z [B,L] with ``carry_state=True``."""
        if self._ctx is None:
            raise _lib.ArcvaeError("decoder.backward() without a forward")
        lib = _lib.load()
        if dlogits is None:                      # after forward_ce: the bf16 d logits are in the tape
            B, T, cond, tape = self._ctx
            scratch = self.ws.get("scratch", lib.arcvae_decoder_scratch_bytes(self._dims, B, T))
            _lib.check(lib.arcvae_decoder_backward(self._dims, self._cparams, cond.data_ptr(), B, T, None,
                                                   tape.data_ptr(), tape.numel(), self._cgrads, scratch.data_ptr(),
                                                   scratch.numel(), self.precision, _lib.stream_ptr()))
            self._ctx = None
            return None
        _lib.require_cuda(dlogits)
        d_tm = dlogits.transpose(0, 1)
        if not d_tm.is_contiguous() or d_tm.dtype != torch.float32:
            d_tm = d_tm.contiguous().float()
        if self.carry_state:
            B, T, cond, tape, zz = self._ctx
            scratch = self.ws.get("scratch_cs", lib.arcvae_decoder_cs_scratch_bytes(self._dims, B, T))
            dz = torch.empty_like(zz)
            _lib.check(lib.arcvae_decoder_cs_backward(self._dims, self._cparams, zz.data_ptr(), cond.data_ptr(), B, T,
                                                      d_tm.data_ptr(), tape.data_ptr(), tape.numel(), self._cgrads,
                                                      dz.data_ptr(), scratch.data_ptr(), scratch.numel(), self.precision,
                                                      _lib.stream_ptr()))
            self._ctx = None
            self.last_dz = dz
            return dz
        B, T, cond, tape = self._ctx
        sbytes = lib.arcvae_decoder_scratch_bytes(self._dims, B, T)
        scratch = self.ws.get("scratch", sbytes)
        _lib.check(lib.arcvae_decoder_backward(self._dims, self._cparams, cond.data_ptr(), B, T, d_tm.data_ptr(),
                                               tape.data_ptr(), tape.numel(), self._cgrads, scratch.data_ptr(),
                                               scratch.numel(), self.precision, _lib.stream_ptr()))
        self._ctx = None
