"""CPU tests (gloo, world_size 2) of the data-parallel protocol in mlx-vae_b200/parallel.py.

The kernels need a GPU; what runs here is the HOST logic: shard ranges, the two-phase loss protocol (per-rank batch
statistics -> all-reduce(sum) -> per-rank gradients scaled by the GLOBAL batch) and the flat-gradient all-reduce.
Each rank evaluates its shard with the CPU oracle's formulas restated on the statistics buffer layout of
include/arcvae_b200.h (ARCVAE_STATS_REDUCE); the combined result must equal the oracle on the full batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _latent_stats(mu, logvar, free_bits):
    """Phase 1 of csrc/loss.cu on one shard: [sum m (L), sum exp(s) (L), sum kl_raw, sum kl_fb, ce, ntok, count]."""
    L = mu.shape[1]
    var = torch.exp(logvar)
    k = -0.5 * (1 + logvar - mu * mu - var)
    kc = torch.clamp(k, min=0.0)
    if free_bits > 0:
        kc = torch.clamp(kc, min=free_bits / L)
    return torch.cat([mu.sum(0), var.sum(0), k.sum().reshape(1), kc.sum().reshape(1), torch.zeros(2, dtype=mu.dtype),
                      torch.tensor([float(mu.shape[0])], dtype=mu.dtype)])


def _latent_grads(mu, logvar, stats, beta, lam_c, lam_mi, free_bits, target=4.85):
    """Phase 2 of csrc/loss.cu: gradients of beta*KL + collapse + mi_penalty w.r.t. this shard's mu / logvar given the
    GLOBAL statistics."""
    L = mu.shape[1]
    Bg = stats[2 * L + 4]
    mbar, vbar = stats[:L] / Bg, stats[L:2 * L] / Bg
    mean_kl = stats[2 * L] / Bg
    agg = -0.5 * (1 + torch.log(vbar) - mbar * mbar - vbar).sum()
    mi_raw = mean_kl - agg
    mi = torch.clamp(mi_raw, min=0.0)
    w = -((lam_c if 4.85 - mi >= 0 else 0.0) + (lam_mi if target - mi >= 0 else 0.0)) * (1.0 if mi_raw > 0 else 0.0)
    var = torch.exp(logvar)
    k = -0.5 * (1 + logvar - mu * mu - var)
    gate = (k > 0).to(mu.dtype)
    if free_bits > 0:
        gate = gate * (torch.clamp(k, min=0.0) > free_bits / L).to(mu.dtype)
    gm = beta / Bg * gate * mu + w / Bg * (mu - mbar)
    gs = beta / Bg * gate * 0.5 * (var - 1) + w / Bg * 0.5 * (var / vbar - 1)
    return gm, gs, float(mi)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("par", os.path.join(ROOT, "mlx-vae_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec); spec.loader.exec_module(par)
    import arcvae_oracle as O
    sync = par.GradSync()
    assert sync.enabled and sync.world == world and sync.rank == rank
    B, L = 37, 8                                   # ragged: 19 + 18
    g = torch.Generator().manual_seed(0)
    mu = torch.tanh(torch.randn(B, L, generator=g, dtype=torch.float64)) * 2
    lv = torch.tanh(torch.randn(B, L, generator=g, dtype=torch.float64)) - 1
    lo, hi = par.shard_range(B, rank, world)
    stats = _latent_stats(mu[lo:hi], lv[lo:hi], 1.0)
    sync.allreduce_stats(stats)
    gm, gs, mi = _latent_grads(mu[lo:hi], lv[lo:hi], stats, 0.05, 0.001, 0.01, 1.0)
    # flat "gradient" all-reduce: each rank contributes its shard's rows, zeros elsewhere -> sum == full tensor
    flat = torch.zeros(2 * B * L, dtype=torch.float64)
    flat[: B * L].view(B, L)[lo:hi] = gm
    flat[B * L:].view(B, L)[lo:hi] = gs
    sync.allreduce_async(flat)
    sync.wait()
    # oracle on the full batch
    m = mu.clone().requires_grad_(True); s = lv.clone().requires_grad_(True)
    total = 0.05 * O.kl_divergence(m, s, free_bits=1.0) + O.posterior_collapse(m, s, weight=0.001) + \
        0.01 * O._mx_maximum(torch.zeros((), dtype=torch.float64), 4.85 - O.mutual_information(m, s))
    total.backward()
    ok = (torch.allclose(flat[: B * L].view(B, L), m.grad, atol=1e-12) and
          torch.allclose(flat[B * L:].view(B, L), s.grad, atol=1e-12) and
          abs(mi - float(O.mutual_information(mu, lv))) < 1e-12)
    q.put((rank, bool(ok), sync.bytes_reduced))
    dist.destroy_process_group()


def test_shard_range_partitions_evenly():
    import importlib.util
    spec = importlib.util.spec_from_file_location("par", os.path.join(ROOT, "mlx-vae_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec); spec.loader.exec_module(par)
    for n in (0, 1, 7, 64, 4096, 32768 + 3):
        for w in (1, 2, 3, 8):
            ranges = [par.shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_protocol_reproduces_global_batch_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(nbytes > 0 for _, _, nbytes in res)
