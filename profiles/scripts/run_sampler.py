#!/usr/bin/env python3
"""One fused-sampler call (for ncu): python profiles/scripts/run_sampler.py [B] [T] [multinomial]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlx_vae_b200 as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
multi = len(sys.argv) > 3 and sys.argv[3] == "1"
dims = dict(vocab_size=80, embedding_dim=128, hidden_dim=256, latent_dim=128, num_conditions=1, num_layers=2)
vae = M.ARCVAE(**dims, seed=67, precision="bf16")
s = M.MLXAutoregressiveDecoderSampling(**dims, decoder=vae.decoder)
c = torch.linspace(-2.5, 2.5, B, device="cuda").unsqueeze(1)
for _ in range(2):
    toks = s.generate_with_temperature(None, c, max_length=T, early_stopping=False, multinomial=multi, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
toks = s.generate_with_temperature(None, c, max_length=T, early_stopping=False, multinomial=multi, seed=1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"B={B} T={T} multinomial={multi}: {ms:.3f} ms, {B / ms * 1e3:.0f} molecules/s, {1e3 * ms / T / ((B + 127) // 128) * min(148, (B + 127) // 128):.2f} us per step per tile-CTA")
