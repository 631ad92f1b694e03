#!/usr/bin/env python3
"""Per-step timeline of the K-split backward cluster kernel (lstm_bwd2_kernel) from globaltimer stamps (debug aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlx_vae_b200 as M
from mlx_vae_b200.data import synthetic_batch
B, T = 4096, 128
x, cond, eps, tf = synthetic_batch(B, T)
enc = M.MLXEncoder(80, 128, 256, 128, 1, 2, seed=1, precision="bf16")
dx, dc = torch.as_tensor(x).cuda(), torch.as_tensor(cond).cuda()
lib = M._lib.load()
buf = torch.zeros(4 * 64 * 32, dtype=torch.int64, device="cuda")
mu, lv = enc(dx, dc)
lib.arcvae_debug_set_rc_stamps(buf.data_ptr()); enc.zero_grad(); enc.backward(torch.ones_like(mu) / B, torch.ones_like(lv) / B)
torch.cuda.synchronize(); lib.arcvae_debug_set_rc_stamps(None)
s = buf.cpu().numpy().reshape(4, 64, 32)
names = {4: "A:start", 5: "A:math+At", 7: "A:arrived", 0: "mma:a_ready", 2: "mma:issued", 8: "st:a_ready", 9: "st:read done", 10: "prefetch issued",
         11: "B:acc_full", 12: "B:tmem+stores", 13: "B:fence", }
order = [4, 5, 7, 0, 8, 10, 2, 9, 11, 12, 13]
for c in range(4):
    print(f"== CTA {c}: ns relative to A:start of the same step")
    for it in range(20, 24):
        base = s[c][it][4]
        print(f"  it={it} step_len={s[c][it + 1][4] - base:6d} | " + " ".join(f"{names[k]}={s[c][it][k] - base:6d}" for k in order if k != 4))
