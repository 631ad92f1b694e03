"""Synthetic SELFIES-shaped batches for benchmarks and smoke runs.

Shape/dtype contract of the reference's data path (mlx_data/dataloader.py:76-82): right-padded with pad_token=0 to a
fixed max_length, end token 2 at the last real position (models/decoder.py:25-26), tokens unsigned 32-bit in the
reference / int32 here, one z-scored property column (TPSA, dataloader.py:63-65)."""
import numpy as np


def synthetic_batch(B: int, T: int, vocab_size: int = 80, num_conditions: int = 1, latent_dim: int = 128,
                    seed: int = 67, tf_ratio: float = 0.9, end_token: int = 2):
    """Lengths U{min(16,T)..T}; body tokens U{3..V-1}; returns (x int32 [B,T], cond f32 [B,C], eps f32 [B,L],
    tf_mask bool [T]) — tf_mask[t] plays the host coin of decoder.py:180."""
    rng = np.random.default_rng(seed)
    lo = min(16, T)
    lens = rng.integers(lo, T + 1, size=B)
    body = rng.integers(3, vocab_size, size=(B, T)).astype(np.int32)
    pos = np.arange(T)[None, :]
    x = np.where(pos < (lens[:, None] - 1), body, 0).astype(np.int32)
    x[np.arange(B), lens - 1] = end_token
    cond = rng.standard_normal((B, num_conditions)).astype(np.float32)
    eps = rng.standard_normal((B, latent_dim)).astype(np.float32)
    tf_mask = rng.random(T) < tf_ratio
    return x, cond, eps, tf_mask
