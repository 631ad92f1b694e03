#!/usr/bin/env python3
"""Stability run of the training step: many steps through the trainer mirror with batch sizes that change the geometry of
the cluster kernels' exchange buffers (stale flags of another geometry must never match), device error flag polled, losses
finite and decreasing on a fixed batch."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlx_vae_b200 as M
from mlx_vae_b200.data import synthetic_batch
enc = M.MLXEncoder(80, 128, 256, 128, 1, 2, seed=1, precision="bf16")
dec = M.MLXAutoregressiveDecoder(80, 128, 256, 128, 1, 2, seed=2, precision="bf16")
tr = M.ARCVAETrainerWithLoss(enc, dec, None, None, learning_rate=2e-4, lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01)
batches = {}
for B in (4096, 200, 130, 1024, 64):
    x, cond, eps, tf = synthetic_batch(B, 128)
    batches[B] = (torch.as_tensor(x).cuda(), torch.as_tensor(cond).cuda())
first, last = {}, {}
n = int(os.environ.get("STRESS_STEPS", 300))
for i in range(n):
    B = (4096, 200, 130, 1024, 64)[i % 5]
    x, c = batches[B]
    out = tr.train_step(x, c, 0.05, 0.9)
    v = float(out["total_loss"])
    assert np.isfinite(v), (i, B, v)
    first.setdefault(B, v); last[B] = v
    if i % 50 == 49:
        tr.check_device_error()
tr.check_device_error()
print("steps", n, {B: (round(first[B], 3), round(last[B], 3)) for B in first})
assert all(last[B] < first[B] for B in first), "loss did not decrease"
print("stress ok")
