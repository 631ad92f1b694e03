#!/usr/bin/env python3
"""Device time of the recurrence launches alone (library timing scopes): encoder forward = layer 0 + layer 1 launches,
backward = layer 1 + layer 0 launches, un-instrumented (no stamps)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlx_vae_b200 as M
from mlx_vae_b200.data import synthetic_batch
B, T = int(os.environ.get("RC_B", 4096)), int(os.environ.get("RC_T", 128))
x, cond, eps, tf = synthetic_batch(B, T)
enc = M.MLXEncoder(80, 128, 256, 128, 1, 2, seed=1, precision="bf16")
dx, dc = torch.as_tensor(x).cuda(), torch.as_tensor(cond).cuda()
mu, lv = enc(dx, dc)
g = torch.ones_like(mu) / B
enc.zero_grad(); enc.backward(g, g); enc.check()
M._lib.timing_enable(True)
n = 5
M._lib.timing_read()
for _ in range(n):
    enc(dx, dc)
torch.cuda.synchronize()
f = M._lib.timing_read()["recurrence"]
for _ in range(n):
    enc(dx, dc); enc.zero_grad(); enc.backward(g, g)
torch.cuda.synchronize()
fb = M._lib.timing_read()["recurrence"]
M._lib.timing_enable(False)
enc.check()
print(f"B={B} T={T}: forward launches {f[0] / n:.3f} ms ({f[1] // n} launches), backward launches {(fb[0] - f[0]) / n:.3f} ms")
