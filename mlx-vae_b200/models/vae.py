"""ARCVAE — drop-in for models/vae.py of the reference."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .decoder import MLXAutoregressiveDecoder
from .decoder_sampling import MLXAutoregressiveDecoderSampling
from .encoder import MLXEncoder


class ARCVAE:
    """models/vae.py:8-131: ``.encoder``, ``.decoder``, ``.decoder_sampling`` (independently initialised, F9), ``.latent_dim``."""

    def __init__(self, vocab_size: int, embedding_dim: int = 256, hidden_dim: int = 512, latent_dim: int = 200,
                 num_conditions: int = 6, num_layers: int = 3, dropout: float = 0.2, *, device=None,
                 seed: Optional[int] = None, precision="fp32", carry_state: bool = False):
        s = (lambda k: None if seed is None else seed + k)
        self.encoder = MLXEncoder(vocab_size, embedding_dim, hidden_dim, latent_dim, num_conditions, num_layers,
                                  dropout, device=device, seed=s(0), precision=precision)
        self.decoder = MLXAutoregressiveDecoder(vocab_size, embedding_dim, hidden_dim, latent_dim, num_conditions,
                                                num_layers, device=device, seed=s(1), precision=precision,
                                                carry_state=carry_state)
        self.decoder_sampling = MLXAutoregressiveDecoderSampling(vocab_size, embedding_dim, hidden_dim, latent_dim,
                                                                 num_conditions, num_layers, device=device, seed=s(2),
                                                                 precision=precision, carry_state=carry_state)
        self.carry_state = bool(carry_state)
        self.latent_dim = latent_dim

    def __call__(self, x: torch.Tensor, conditions: torch.Tensor, target_seq: Optional[torch.Tensor] = None,
                 teacher_forcing_ratio: float = 0.5, *, eps: Optional[torch.Tensor] = None, tf_mask=None,
                 seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """vae.py:63-99 -> (recon_logits [B,T,V], mu, logvar, z)."""
        mu, logvar = self.encoder(x, conditions)
        z = self.encoder.reparameterize(mu, logvar, eps, seed=seed)
        logits = self.decoder(z, conditions, target_seq=target_seq, teacher_forcing_ratio=teacher_forcing_ratio,
                              tf_mask=tf_mask)
        return logits, mu, logvar, z

    def generate(self, batch_size: int, conditions: torch.Tensor, max_length: int = 80, temperature: float = 1.0, *,
                 multinomial: bool = False, seed: int = 0, use_trained_decoder: bool = False) -> torch.Tensor:
        """vae.py:101-131.  The prior draw z ~ N(0,I) (:121) is dead (F1) and therefore skipped.
        ``use_trained_decoder`` samples from ``self.decoder`` instead of the never-trained ``decoder_sampling.decoder``."""
        if use_trained_decoder:
            sampler = MLXAutoregressiveDecoderSampling(
                self.decoder.vocab_size, self.decoder.embedding_dim, self.decoder.hidden_dim, self.decoder.latent_dim,
                self.decoder.num_conditions, self.decoder.num_layers, self.decoder.pad_token, self.decoder.end_token,
                decoder=self.decoder)
        else:
            sampler = self.decoder_sampling
        z = None
        if self.carry_state:                     # vae.py:121: z ~ N(0, I) from the prior (Philox(seed) on the device)
            zero = torch.zeros((batch_size, self.latent_dim), dtype=torch.float32, device=self.encoder.device)
            z = MLXEncoder.reparameterize(zero, zero, None, seed=seed)
        return sampler.generate_with_temperature(z, conditions, max_length=max_length, temperature=temperature,
                                                 multinomial=multinomial, seed=seed)
