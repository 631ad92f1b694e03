"""Data parallelism on real hardware: 2 NCCL ranks (one process per GPU) against the single-GPU step on the concatenated
batch.  Skipped on boxes with fewer than 2 GPUs (the driver's multi-GPU tier and `gpurun --gpus 2` run it).

What must hold: every scalar of the loss dict, every gradient and the weights
after the Adam update are those of the GLOBAL batch — including recon_loss / total_loss, which round 1 reported divided
by the world size."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

DIMS_SMALL = dict(vocab_size=23, embedding_dim=16, hidden_dim=32, latent_dim=16, num_conditions=1, num_layers=2)
DIMS_DEFAULT = dict(vocab_size=80, embedding_dim=128, hidden_dim=256, latent_dim=128, num_conditions=1, num_layers=2)
HYPER = dict(lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01)


def _step(M, dims, precision, x, cond, eps, tf_mask, group, seed=None):
    vae = M.ARCVAE(**dims, seed=11, precision=precision)
    tr = M.ARCVAETrainerWithLoss(vae.encoder, vae.decoder, None, None, learning_rate=1e-3, batch_size=x.shape[0],
                                 process_group=group, **HYPER)
    d = tr.train_step(x, cond, 0.05, 0.9, eps=eps, tf_mask=tf_mask, seed=seed)
    torch.cuda.synchronize()
    tr.check_device_error()
    return vae, tr, d


def _worker(rank, world, port, dims, precision, B, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import mlx_vae_b200 as M
    from mlx_vae_b200.data import synthetic_batch
    x, cond, eps, tf_mask = synthetic_batch(B, T, vocab_size=dims["vocab_size"], latent_dim=dims["latent_dim"], seed=5)
    lo, hi = M.parallel.shard_range(B, rank, world)
    dx, dc, de = (torch.as_tensor(a[lo:hi]).cuda() for a in (x, cond, eps))
    # (a) injected eps
    vae, tr, d = _step(M, dims, precision, dx, dc, de, tf_mask, None if world == 1 else dist.group.WORLD)
    out = {"loss": {k: float(d[k]) for k in M._lib.LOSS_KEYS},
           "genc": vae.encoder.grads.flat.cpu().numpy(), "gdec": vae.decoder.grads.flat.cpu().numpy(),
           "penc": vae.encoder.params.flat.cpu().numpy(), "pdec": vae.decoder.params.flat.cpu().numpy()}
    # (b) Philox eps: the draw is indexed by the GLOBAL row, so the shards reproduce the single-device z
    _, _, d2 = _step(M, dims, precision, dx, dc, None, tf_mask, None if world == 1 else dist.group.WORLD, seed=9)
    out["z"] = d2["z"].cpu().numpy()
    out["lo_hi"] = (lo, hi)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _single(dims, precision, B, T):
    import mlx_vae_b200 as M
    from mlx_vae_b200.data import synthetic_batch
    x, cond, eps, tf_mask = synthetic_batch(B, T, vocab_size=dims["vocab_size"], latent_dim=dims["latent_dim"], seed=5)
    dx, dc, de = (torch.as_tensor(a).cuda() for a in (x, cond, eps))
    vae, tr, d = _step(M, dims, precision, dx, dc, de, tf_mask, None)
    out = {"loss": {k: float(d[k]) for k in M._lib.LOSS_KEYS},
           "genc": vae.encoder.grads.flat.cpu().numpy(), "gdec": vae.decoder.grads.flat.cpu().numpy(),
           "penc": vae.encoder.params.flat.cpu().numpy(), "pdec": vae.decoder.params.flat.cpu().numpy()}
    _, _, d2 = _step(M, dims, precision, dx, dc, None, tf_mask, None, seed=9)
    out["z"] = d2["z"].cpu().numpy()
    return out


def _rel(a, b):
    return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-30)


@pytest.mark.parametrize("dims,precision,B,T,tol", [(DIMS_SMALL, "fp32", 48, 9, 1e-5), (DIMS_DEFAULT, "fp32", 64, 12, 1e-5),
                                                   (DIMS_DEFAULT, "bf16", 256, 16, 2e-2)])
def test_two_ranks_equal_single_gpu_global_batch(dims, precision, B, T, tol):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ref = _single(dims, precision, B, T)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, precision, B, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank in (0, 1):
        o = res[rank]
        for k, v in ref["loss"].items():          # the loss dict is the GLOBAL batch's on every rank
            assert abs(o["loss"][k] - v) <= tol * max(1.0, abs(v)), (rank, k, o["loss"][k], v)
        for k in ("genc", "gdec"):
            assert _rel(o[k], ref[k]) <= (tol if precision == "fp32" else 5e-2), (rank, k, _rel(o[k], ref[k]))
        for k in ("penc", "pdec"):                # one Adam step of 1e-3 * O(3) on weights of O(0.2)
            assert _rel(o[k], ref[k]) <= (1e-4 if precision == "fp32" else 2e-2), (rank, k, _rel(o[k], ref[k]))
        lo, hi = o["lo_hi"]
        # Philox eps is indexed by the GLOBAL row: the shard's z equals the single-device z up to the rounding of mu/logvar
        # (a wrong offset gives unrelated N(0,1) draws, differences of O(1))
        assert np.abs(o["z"] - ref["z"][lo:hi]).max() < (1e-4 if precision == "fp32" else 5e-2)
    # the replicas hold identical weights after the step
    assert np.array_equal(res[0]["penc"], res[1]["penc"]) and np.array_equal(res[0]["pdec"], res[1]["pdec"])


def test_replicas_start_from_rank0_parameters():
    """seed=None initialises every rank differently; the trainer broadcasts rank 0's parameters."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert not np.array_equal(res[0][0], res[1][0]), "different seeds should give different initial weights"
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][1], res[0][0])


def _bcast_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import mlx_vae_b200 as M
    vae = M.ARCVAE(**DIMS_SMALL, seed=None)
    before = vae.encoder.params.flat.cpu().numpy()
    M.ARCVAETrainerWithLoss(vae.encoder, vae.decoder, None, None, process_group=dist.group.WORLD)
    torch.cuda.synchronize()
    q.put((rank, (before, vae.encoder.params.flat.cpu().numpy())))
    dist.barrier()
    dist.destroy_process_group()


def test_error_flag_is_sticky_and_guards_adam():
    """A timed-out persistent kernel raises a sticky device flag: Adam must not touch the weights and the trainer raises."""
    import mlx_vae_b200 as M
    lib = M._lib.load()
    vae = M.ARCVAE(**DIMS_SMALL, seed=3)
    tr = M.ARCVAETrainerWithLoss(vae.encoder, vae.decoder, None, None, learning_rate=1e-2, batch_size=8, error_check_every=1, **HYPER)
    x = torch.randint(0, 23, (8, 7), device="cuda"); c = torch.randn(8, 1, device="cuda")
    tf_mask = np.ones(7, dtype=bool)
    tr.train_step(x, c, 0.05, 0.9, tf_mask=tf_mask)
    p0 = vae.encoder.params.flat.clone()
    try:
        M._lib.check(lib.arcvae_debug_raise_device_error(M._lib.stream_ptr()))
        with pytest.raises(M._lib.ArcvaeError):
            tr.train_step(x, c, 0.05, 0.9, tf_mask=tf_mask)
        assert torch.equal(vae.encoder.params.flat, p0), "Adam applied gradients while the error flag was set"
        vae.encoder(x, c)                                   # a new forward must NOT clear the flag
        assert M._lib.device_error_flag() != 0
    finally:
        M._lib.clear_device_error()
    assert M._lib.device_error_flag() == 0
    tr.train_step(x, c, 0.05, 0.9, tf_mask=tf_mask)
    assert not torch.equal(vae.encoder.params.flat, p0)
