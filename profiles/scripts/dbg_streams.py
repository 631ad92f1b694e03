import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import mlx_vae_b200 as M
from mlx_vae_b200.data import synthetic_batch
DIMS = dict(vocab_size=80, embedding_dim=128, hidden_dim=256, latent_dim=128, num_conditions=1, num_layers=2)
x, cond, eps, tf_mask = synthetic_batch(4096, 128, seed=67, tf_ratio=0.9)
vae = M.ARCVAE(**DIMS, seed=67, precision="bf16")
tr = M.ARCVAETrainerWithLoss(vae.encoder, vae.decoder, None, None, learning_rate=2e-4, batch_size=4096, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01)
dx, dc, de = (torch.as_tensor(a).cuda() for a in (x, cond, eps))
def run(tag, n=10):
    for _ in range(3): tr.train_step(dx, dc, 0.05, 0.9, eps=de, tf_mask=tf_mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): tr.train_step(dx, dc, 0.05, 0.9, eps=de, tf_mask=tf_mask)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(tag, "gpu ms/step", e0.elapsed_time(e1) / n, "cpu enqueue ms/step", 1e3 * (t1 - t0) / n, "wall", 1e3 * (t2 - t0) / n)
run("two streams")
os.environ["ARCVAE_ONE_STREAM"] = "1"
run("serial after two-stream")
run("serial again")
os.environ.pop("ARCVAE_ONE_STREAM")
run("two streams again")
