"""MLXAutoregressiveDecoderSampling — drop-in for models/decoder_sampling.py of the reference."""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib
from .decoder import MLXAutoregressiveDecoder


class MLXAutoregressiveDecoderSampling:
    """Sampler wrapper (models/decoder_sampling.py:6-128).  Like the reference it owns a SEPARATE decoder
    (``self.decoder``, :29-38, F9); pass ``decoder=`` to sample from trained weights instead."""

    def __init__(self, vocab_size: int, embedding_dim: int = 256, hidden_dim: int = 512, latent_dim: int = 200,
                 num_conditions: int = 6, num_layers: int = 3, pad_token: int = 0, end_token: int = 2, *,
                 decoder: Optional[MLXAutoregressiveDecoder] = None, device=None, seed: Optional[int] = None,
                 precision="fp32", carry_state: bool = False):
        self.decoder = decoder if decoder is not None else MLXAutoregressiveDecoder(
            vocab_size, embedding_dim, hidden_dim, latent_dim, num_conditions, num_layers, pad_token, end_token,
            device=device, seed=seed, precision=precision, carry_state=carry_state)
        self.vocab_size, self.embedding_dim, self.hidden_dim = vocab_size, embedding_dim, hidden_dim
        self.latent_dim, self.num_conditions = latent_dim, num_conditions
        self.pad_token, self.end_token = pad_token, end_token

    def generate_with_temperature(self, z: torch.Tensor, conditions: torch.Tensor, max_length: int = 80,
                                  temperature: float = 1.0, early_stopping: bool = True, *,
                                  multinomial: bool = False, seed: int = 0) -> torch.Tensor:
        """decoder_sampling.py:48-128.  Returns tokens [B, t_stop] int32 with t_stop <= max_length: generation stops
        before the first step at which every row has emitted ``end_token`` (:87-88); rows are not padded after their
        own end (F8).  ``multinomial=False`` is the reference's argmax(softmax(logits/temperature)) (:110-117);
        ``multinomial=True`` draws from the categorical with Philox(seed) — the reference's TODO (:116).
        ``z`` is accepted and unused, as in the reference (F1).  Reading t_stop synchronises the stream once."""
        lib = _lib.load()
        dec = self.decoder
        cond = dec._f32(conditions)
        B = cond.shape[0]
        if getattr(dec, "carry_state", False):
            return self._generate_with_state(z, cond, int(max_length), early_stopping, multinomial)
        nbytes = lib.arcvae_sampler_workspace_bytes(dec._dims, B, max_length)
        ws = dec.ws.get("sampler", nbytes)
        tokens = torch.empty((B, max(max_length, 1)), dtype=torch.int32, device=dec.device)
        t_stop = torch.empty((1,), dtype=torch.int32, device=dec.device)
        _lib.check(lib.arcvae_sample(dec._dims, dec._cparams, cond.data_ptr(), B, int(max_length), float(temperature),
                                     1 if early_stopping else 0, 1 if multinomial else 0, int(seed), tokens.data_ptr(),
                                     t_stop.data_ptr(), ws.data_ptr(), ws.numel(), dec.precision, _lib.stream_ptr()))
        n = int(t_stop.item())
        return tokens[:, :n]

    def _generate_with_state(self, z, cond, max_length: int, early_stopping: bool, multinomial: bool) -> torch.Tensor:
        """carry_state=True (extension): greedy chain with the LSTM state initialised from z and carried.
        The chain is the decoder's own feedback path (decoder.py:185) run for max_length + 1 positions: the token fed at
        position t+1 IS the token emitted at t.  Greedy only; temperature does not change an argmax."""
        if multinomial:
            raise NotImplementedError("carry_state=True sampler: greedy (the reference's argmax path) only")
        if max_length <= 0:
            return torch.empty((cond.shape[0], 0), dtype=torch.int32, device=cond.device)
        dec = self.decoder
        dec(z, cond, target_seq=None, max_length=max_length + 1)
        tokens = dec.last_inputs[:, 1:].contiguous()                      # [B, max_length]
        if not early_stopping:
            return tokens
        # the reference stops BEFORE the first step at which every row has already emitted end_token (:87-88)
        ended = (tokens == dec.end_token).to(torch.int32).cumsum(dim=1) > 0
        all_ended = ended.all(dim=0)
        idx = torch.nonzero(all_ended)
        n = int(idx[0]) + 1 if idx.numel() else max_length
        return tokens[:, :n]
