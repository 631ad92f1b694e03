#!/usr/bin/env python3
"""One stand-alone launch of the fused loss kernel on the benched shapes (logits [T,B,V] fp32, d logits in place): the
kernel of complete_vae_loss (evaluation) and of the fp32 mode; on the bf16 training path CE runs in the fc_out epilogue."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlx_vae_b200 as M  # noqa: E402,F401
from mlx_vae_b200.losses._fused import fused_loss, make_hyper  # noqa: E402

B, T, V, L = 4096, 128, 80, 128
lg = torch.randn((T, B, V), device="cuda").transpose(0, 1)
x = torch.randint(0, V, (B, T), device="cuda", dtype=torch.int32)
mu = torch.tanh(torch.randn(B, L, device="cuda")); lv = torch.tanh(torch.randn(B, L, device="cuda")) - 1
eps = torch.randn(B, L, device="cuda")
hp = make_hyper(0.05, 0.1, 0.001, 1.0, 0.01)
for _ in range(2):
    fused_loss(lg, x, mu, lv, hp, eps=eps, inplace_dlogits=True)
torch.cuda.synchronize()
print("ok")
