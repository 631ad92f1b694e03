import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mlx_vae_b200 as M
B, T = 148 * 128, 64
for H, NL in ((256, 2), (256, 1), (256, 3), (128, 2), (64, 2)):
    dims = dict(vocab_size=80, embedding_dim=128, hidden_dim=H, latent_dim=128, num_conditions=1, num_layers=NL)
    vae = M.ARCVAE(**dims, seed=67, precision="bf16")
    s = M.MLXAutoregressiveDecoderSampling(**dims, decoder=vae.decoder)
    c = torch.linspace(-2.5, 2.5, B, device="cuda").unsqueeze(1)
    for _ in range(2):
        s.generate_with_temperature(None, c, max_length=T, early_stopping=False, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.generate_with_temperature(None, c, max_length=T, early_stopping=False, seed=1); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"H={H} NL={NL}: {1e3 * ms / T:.2f} us per step per tile")
