#!/usr/bin/env python3
"""Run-to-run determinism of the encoder forward + BPTT (cluster recurrence): max |difference| of mu/logvar and of every
gradient between repeated runs on identical inputs, relative to the tensor's max magnitude (debug aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mlx_vae_b200 as M
import arcvae_oracle as O
from mlx_vae_b200.data import synthetic_batch
for B, T in ((128, 2), (128, 9), (200, 16), (4096, 128)):
    x, cond, eps, tf = synthetic_batch(B, T)
    enc = M.MLXEncoder(80, 128, 256, 128, 1, 2, seed=1, precision="bf16")
    dx, dc = torch.as_tensor(x).cuda(), torch.as_tensor(cond).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    dmu = torch.randn(B, 128, device="cuda", generator=g) / B
    dlv = torch.randn(B, 128, device="cuda", generator=g) / B
    ref = None
    worst = {}
    for rep in range(6):
        mu, lv = enc(dx, dc)
        enc.zero_grad(); enc.backward(dmu, dlv); enc.check()
        cur = {"mu": mu.clone(), "lv": lv.clone()}
        cur.update({k: v.clone() for k, v in O.tree_flatten(enc.gradients()).items()})
        if ref is None:
            ref = cur
            continue
        for k in ref:
            s = float(ref[k].abs().max())
            if s > 0:
                worst[k] = max(worst.get(k, 0.0), float((cur[k] - ref[k]).abs().max()) / s)
    bad = {k: f"{v:.1e}" for k, v in worst.items() if v > 1e-5}
    print(f"B={B} T={T}: max rel run-to-run difference {max(worst.values()):.2e}; tensors above 1e-5: {bad}")
