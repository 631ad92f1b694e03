"""Test infrastructure: pure-Python Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; Random123) and the N(0,1) / uniform transforms the kernels use (csrc/kernels.cuh philox4x32, u01;
csrc/loss.cu philox_normal).  The reference draws from MLX's unseeded global threefry key (encoder.py:150), which is
not reproducible, so the device RNG has no reference stream to match — only this restatement."""
import math

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32(seed: int, ctr_lo: int, ctr_hi: int):
    c = [ctr_lo & MASK, (ctr_lo >> 32) & MASK, ctr_hi & MASK, (ctr_hi >> 32) & MASK]
    k = [seed & MASK, (seed >> 32) & MASK]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return tuple(c)


def u01(x: int) -> np.float32:
    return np.float32((np.float32(x >> 8) + np.float32(1.0)) * np.float32(1.0 / 16777216.0))


def normal(seed: int, ctr: int) -> float:
    r = philox4x32(seed, ctr, 0)
    u1, u2 = float(u01(r[0])), float(u01(r[1]))
    return math.sqrt(-2.0 * math.log(u1)) * math.cos(6.283185307179586 * u2)
