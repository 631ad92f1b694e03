#!/usr/bin/env python3
"""(needs make -C mlx-vae_b200/csrc clean all STAMPS=1)  Timeline of ONE sampler step (t = 10, CTA 0): per 192-column block
the MMA thread's start / commit and the cell-math thread's start / end, plus the selection."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mlx_vae_b200 as M
dims = dict(vocab_size=80, embedding_dim=128, hidden_dim=256, latent_dim=128, num_conditions=1, num_layers=2)
vae = M.ARCVAE(**dims, seed=67, precision="bf16")
s = M.MLXAutoregressiveDecoderSampling(**dims, decoder=vae.decoder)
B = 148 * 128
c = torch.linspace(-2.5, 2.5, B, device="cuda").unsqueeze(1)
s.generate_with_temperature(None, c, max_length=32, early_stopping=False, seed=1)
buf = torch.zeros(4 * 64 * 32, dtype=torch.int64, device="cuda")
lib = M._lib.load()
lib.arcvae_debug_set_rc_stamps(buf.data_ptr())
s.generate_with_temperature(None, c, max_length=32, early_stopping=False, seed=1)
torch.cuda.synchronize(); lib.arcvae_debug_set_rc_stamps(None)
v = buf.cpu().numpy()
base = v[0]
for l in range(2):
    for j in range(4):
        o = (l * 4 + j) * 4
        print(f"layer {l} block {j}: mma start {v[o] - base:6d}  mma committed {v[o + 1] - base:6d} | cell start {v[o + 2] - base:6d}  cell end {v[o + 3] - base:6d}")
print(f"selection start {v[60] - base:6d}, one-hot written {v[61] - base:6d}")
