#!/usr/bin/env python3
"""Per-phase device time of the scaled-config encoder (H=1024, 3 layers): forward / backward per timestep, with and
without the 256-row tiles.  python profiles/scripts/scaled_rec_probe.py [B] [T]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlx_vae_b200 as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
kw = dict(vocab_size=80, embedding_dim=128, hidden_dim=1024, latent_dim=256, num_conditions=1, num_layers=3)


def run(tag):
    enc = M.MLXEncoder(**kw, seed=1, precision="bf16")
    x = torch.randint(3, 80, (B, T), device="cuda", dtype=torch.int32)
    c = torch.randn(B, 1, device="cuda")
    dmu = torch.randn(B, 256, device="cuda") / B
    for _ in range(2):
        mu, lv = enc(x, c); enc.zero_grad(); enc.backward(dmu, dmu)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    M._lib.timing_enable(True); M._lib.timing_read()
    ev[0].record(); mu, lv = enc(x, c); ev[1].record(); enc.zero_grad(); enc.backward(dmu, dmu); ev[2].record()
    torch.cuda.synchronize()
    cats = M._lib.timing_read(); M._lib.timing_enable(False)
    f, b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(f"{tag}: B={B} T={T} forward {f:.2f} ms ({1e3 * f / (3 * T):.1f} us per layer-step), backward {b:.2f} ms "
          f"({1e3 * b / (3 * T):.1f} us per layer-step); categories {dict((k, round(v[0], 2)) for k, v in cats.items() if v[1])}")


run("fused step, 256-row tiles")
os.environ["ARCVAE_NO_BM256"] = "1"
run("fused step, 128-row tiles")
os.environ.pop("ARCVAE_NO_BM256")
os.environ["ARCVAE_NO_FUSED_STEP"] = "1"
run("unfused per-step path (round 1)")
