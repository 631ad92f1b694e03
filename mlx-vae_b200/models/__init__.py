"""Mirrors models/__init__.py of the reference."""
from .decoder import MLXAutoregressiveDecoder
from .decoder_sampling import MLXAutoregressiveDecoderSampling
from .encoder import MLXEncoder
from .vae import ARCVAE

__all__ = ["MLXEncoder", "MLXAutoregressiveDecoder", "MLXAutoregressiveDecoderSampling", "ARCVAE"]
