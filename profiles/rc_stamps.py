#!/usr/bin/env python3
"""Per-step cycle breakdown of the cluster recurrence kernel (CTA 0), from clock64 stamps (debug aid)."""
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
for mode in ("fwd", "bwd"):
    buf = torch.zeros(4 * 64 * 32, dtype=torch.int64, device="cuda")
    mu, lv = enc(dx, dc)
    if mode == "fwd":
        lib.arcvae_debug_set_rc_stamps(buf.data_ptr()); mu, lv = enc(dx, dc)
    else:
        lib.arcvae_debug_set_rc_stamps(buf.data_ptr()); enc.zero_grad(); enc.backward(torch.ones_like(mu) / B, torch.ones_like(lv) / B)
    torch.cuda.synchronize(); lib.arcvae_debug_set_rc_stamps(None)
    sall = buf.cpu().numpy().reshape(4, 64, 32)   # last launch (layer 1 fwd / layer 0 bwd) wins; globaltimer ns
    print(f"== {mode}: per-CTA stamps (ns) relative to CTA0 'acc ready', it=21")
    b0 = sall[0][21][4]
    for c in range(4):
        print(f"  cta{c}: acc_ready={sall[c][21][4]-b0:6d} math_done={sall[c][21][5]-b0:6d} fence={sall[c][21][6]-b0:6d} arrived={sall[c][21][7]-b0:6d} "
              f"snd_seen={sall[c][21][8]-b0:6d} tma_issued={sall[c][21][9]-b0:6d} | next first={sall[c][22][0]-b0:6d} last={sall[c][22][1]-b0:6d} mma_issued={sall[c][22][2]-b0:6d} acc_ready={sall[c][22][4]-b0:6d}")
    s = sall[0]
    names = {0: "ctl:first chunk landed", 1: "ctl:last chunk landed", 2: "ctl:MMAs issued", 4: "epi:acc ready", 5: "epi:math+stores done",
             6: "epi:fence done", 7: "epi:arrived", 8: "snd:epi_done seen", 9: "snd:TMA issued"}
    print(f"== {mode}: cycles relative to 'epi:acc ready' of the same step (it = 20..23)")
    for it in range(20, 24):
        base = s[it][4]
        print(f"it={it} step_len={s[it+1][4]-s[it][4]:6d} | " + " ".join(f"{names[k].split(':')[1][:14]}={s[it][k]-base:6d}" for k in (5, 6, 7, 8, 9)) +
              " || next: " + " ".join(f"{names[k].split(':')[1][:14]}={s[it+1][k]-base:6d}" for k in (0, 1, 2, 4)))
