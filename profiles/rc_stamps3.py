#!/usr/bin/env python3
"""(needs a stamp-enabled library: make -C mlx-vae_b200/csrc clean all STAMPS=1)
Per-step timelines (globaltimer ns) of the flag-in-data cluster recurrence kernels lstm_fwd3_kernel / lstm_bwd3_kernel
and launch times of the four recurrence launches (debug aid; the stamping thread itself runs ~1 us behind the other warps — read per-warp slots, not thread 0)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlx_vae_b200 as M
from mlx_vae_b200.data import synthetic_batch
B, T = int(os.environ.get("RC_B", 4096)), int(os.environ.get("RC_T", 128))
x, cond, eps, tf = synthetic_batch(B, T)
enc = M.MLXEncoder(80, 128, 256, 128, 1, 2, seed=1, precision="bf16")
dx, dc = torch.as_tensor(x).cuda(), torch.as_tensor(cond).cuda()
lib = M._lib.load()
gen = "3"
mu, lv = enc(dx, dc)
enc.zero_grad(); enc.backward(torch.ones_like(mu) / B, torch.ones_like(lv) / B)
enc.check()
# whole-pass timings (forward = 2 recurrence launches + projection + head, backward = 2 launches + GEMMs)
tf_, tb_ = [], []
for _ in range(5):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); enc(dx, dc); e1.record(); enc.zero_grad(); enc.backward(torch.ones_like(mu) / B, torch.ones_like(lv) / B); e2.record()
    torch.cuda.synchronize()
    tf_.append(e0.elapsed_time(e1)); tb_.append(e1.elapsed_time(e2))
print(f"gen {gen} B={B} T={T}: encoder forward pass {min(tf_):.3f} ms, backward pass {min(tb_):.3f} ms (min of 5)")
enc.check()
FWD = {4: "acc_full", 5: "math", 7: "own+LL sent", 10: "prefetch issued", 11: "gathered", 12: "tape done", 0: "mma:own", 1: "mma:last", 2: "mma:issued", 8: "st:seen", 9: "st:read"}
BWD = {4: "A:start", 6: "A:polled", 5: "A:half0", 7: "A:half1", 0: "mma:h0", 3: "mma:h0 issued", 1: "mma:h1", 2: "mma:issued", 8: "st:seen", 9: "st:read", 10: "prefetch",
       11: "B:acc_full", 12: "B:sent"}
for mode, names in (("fwd", FWD), ("bwd", BWD)):
    buf = torch.zeros(4 * 64 * 32, dtype=torch.int64, device="cuda")
    mu, lv = enc(dx, dc)
    lib.arcvae_debug_set_rc_stamps(buf.data_ptr())
    if mode == "fwd":
        mu, lv = enc(dx, dc)
    else:
        enc.zero_grad(); enc.backward(torch.ones_like(mu) / B, torch.ones_like(lv) / B)
    torch.cuda.synchronize(); lib.arcvae_debug_set_rc_stamps(None)
    s = buf.cpu().numpy().reshape(4, 64, 32)   # the last launch wins (layer 1 forward / layer 0 backward)
    order = sorted(names, key=lambda k: np.median([s[0][it][k] - s[0][it][4] for it in range(20, 40)]))
    for c in range(2):
        print(f"== {mode} gen {gen} CTA {c}: ns relative to slot 4 of the same step; step = slot4[it+1]-slot4[it]")
        for it in (20, 21, 40):
            base = s[c][it][4]
            print(f"  it={it} step={s[c][it + 1][4] - base:6d} | " + " ".join(f"{names[k]}={s[c][it][k] - base:6d}" for k in order if k != 4))
    for c in range(4):
        for it in (21, 40):
            base = s[c][it][4]
            if mode == "fwd":
                print(f"  CTA {c} it={it} acc_full vs CTA 0: {s[c][it][4] - s[0][it][4]:5d} | per-warp sent " + " ".join(f"{s[c][it][24 + w] - base:5d}" for w in range(8)))
            print(f"  CTA {c} it={it} per-warp: " + ("gathered " if mode == "fwd" else "polled ") + " ".join(f"{s[c][it][16 + w] - base:5d}" for w in range(8)) +
                  ("" if mode == "fwd" else " | half1 " + " ".join(f"{s[c][it][24 + w] - base:5d}" for w in range(8))))
    if mode == "fwd":
        print("  poll rounds of thread 0 (3 sources, minimum 3): " + " ".join(str(int(s[0][it][13])) for it in range(20, 40)))
    steps = [s[0][it + 1][4] - s[0][it][4] for it in range(8, 62)]
    print(f"== {mode} gen {gen}: median step {np.median(steps):.0f} ns, mean {np.mean(steps):.0f} ns")
