"""GPU tests of the epoch-level evaluation mirrors (trainer.py:116-175 `_compute_true_train_loss`, :418-487 `_validate`;
SURVEY §8f row N3) and of the device-side dataset feeding the trainer (row N2), against the oracle."""
import numpy as np
import pytest
import torch

import arcvae_oracle as O
import dataset_oracle as DO
from _util import model_kwargs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mlx_vae_b200
    return mlx_vae_b200


def test_validate_matches_oracle_and_dataset_feeds_training(M):
    cfg = O.Config(23, 8, 16, 8, 1, 2)
    p = O.init_params(cfg, seed=8, dtype=torch.float64)
    p["decoder"] = O.tree_map(lambda t: t * 3.0, p["decoder"])
    rng = np.random.default_rng(2)
    n, T, bs = 21, 9, 8
    mols = [list(rng.integers(3, cfg.vocab_size, size=int(rng.integers(2, 12))).tolist()) + [2] for _ in range(n)]
    tpsa = rng.normal(70, 20, size=(n, 1))
    kw = model_kwargs(cfg)
    enc = M.MLXEncoder(**kw).load_parameters(p["encoder"])
    dec = M.MLXAutoregressiveDecoder(**kw).load_parameters(p["decoder"])
    ds = M.MoleculeDataset(mols, tpsa, max_length=T)
    ref_ds = DO.MoleculeDatasetOracle(mols, tpsa, max_length=T)
    hyper = dict(lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01)
    tr = M.ARCVAETrainerWithLoss(enc, dec, None, ds, learning_rate=1e-3, batch_size=bs, beta_start=0.0, beta_end=0.05,
                                 beta_warmup_epochs=20, **hyper)
    beta = 0.03
    # oracle: every batch, teacher forcing off (all coins false), mean of the per-batch terms
    acc = {k: 0.0 for k in ("total_loss", "recon_loss", "kl_loss", "collapse_penalty", "prop_loss")}
    nb = 0
    for mb, pb in ref_ds.to_batches(bs, shuffle=False):
        x = torch.as_tensor(mb.astype(np.int64)); c = torch.as_tensor(pb).double()
        d = O.complete_vae_loss(p, x, c, cfg.num_layers, torch.zeros(x.shape[0], cfg.latent_dim, dtype=torch.float64),
                                np.zeros(T, dtype=bool), beta=beta, target_mi=4.85, **hyper)
        for k in acc:
            acc[k] += float(d[k])
        nb += 1
    ref = {"loss": acc["total_loss"] / nb, "recon": acc["recon_loss"] / nb, "kl": acc["kl_loss"] / nb,
           "collapse": acc["collapse_penalty"] / nb, "prop": acc["prop_loss"] / nb}
    got = tr._validate(ds, beta)
    assert set(got) == set(ref)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 1e-4 * max(1.0, abs(ref[k])), (k, got[k], ref[k])
    # _compute_true_train_loss: first num_batches unshuffled batches at beta(epoch)
    got2 = tr._compute_true_train_loss(epoch=12, num_batches=2)
    acc2, nb2 = 0.0, 0
    for mb, pb in ref_ds.to_batches(bs, shuffle=False):
        if nb2 >= 2:
            break
        x = torch.as_tensor(mb.astype(np.int64)); c = torch.as_tensor(pb).double()
        d = O.complete_vae_loss(p, x, c, cfg.num_layers, torch.zeros(x.shape[0], cfg.latent_dim, dtype=torch.float64),
                                np.zeros(T, dtype=bool), beta=tr.compute_beta(12), target_mi=4.85, **hyper)
        acc2 += float(d["total_loss"]); nb2 += 1
    assert abs(got2["loss"] - acc2 / nb2) <= 1e-4 * max(1.0, abs(acc2 / nb2))
    # one epoch of training straight from the device-resident dataset (ragged last batch of 5 included)
    np.random.seed(3)
    out = tr._train_epoch_batches(beta, 0.9)
    assert np.isfinite(out["loss"]) and tr.step_count == 3


def test_checkpoint_roundtrip(M, tmp_path):
    cfg = O.Config(23, 8, 16, 8, 1, 2)
    kw = model_kwargs(cfg)
    enc = M.MLXEncoder(**kw, seed=1); dec = M.MLXAutoregressiveDecoder(**kw, seed=2)
    tr = M.ARCVAETrainerWithLoss(enc, dec, None, None, learning_rate=1e-3, batch_size=8)
    x = torch.randint(0, 23, (8, 7), device="cuda"); c = torch.randn(8, 1, device="cuda")
    tr.train_step(x, c, 0.05, 0.9, tf_mask=np.ones(7, dtype=bool))
    tr.checkpoint_dir = str(tmp_path / "ckdir")
    tr.history["epoch"].append(3); tr.history["train_loss"].append(1.25)
    path = tr.save_checkpoint(3, is_best=True)            # reference call surface (trainer.py:577)
    import os
    assert path.endswith("checkpoint_epoch_003.npz") and os.path.exists(os.path.join(tr.checkpoint_dir, "checkpoint_best.npz"))
    ck = np.load(path)
    assert "encoder/lstm_layer_0.Wh" in ck.files and "decoder/fc_out.weight" in ck.files and "encoder_opt/m/fc_mu.weight" in ck.files
    enc2 = M.MLXEncoder(**kw, seed=7); dec2 = M.MLXAutoregressiveDecoder(**kw, seed=8)
    tr2 = M.ARCVAETrainerWithLoss(enc2, dec2, None, None, learning_rate=1e-3, batch_size=8)
    assert tr2.load_checkpoint(path) == 3
    assert tr2.history["epoch"] == [3] and tr2.history["train_loss"] == [1.25]
    assert tr.save_checkpoint_to(tmp_path / "explicit", epoch=4).endswith("explicit.npz")
    assert torch.equal(enc2.params.flat, enc.params.flat) and torch.equal(dec2.params.flat, dec.params.flat)
    assert torch.equal(tr2.encoder_optimizer.m, tr.encoder_optimizer.m) and torch.equal(tr2.decoder_optimizer.v, tr.decoder_optimizer.v)
    a = tr.train_step(x, c, 0.05, 0.9, tf_mask=np.ones(7, dtype=bool), seed=5)
    b = tr2.train_step(x, c, 0.05, 0.9, tf_mask=np.ones(7, dtype=bool), seed=5)
    assert float(a["total_loss"]) == float(b["total_loss"]) and torch.equal(enc2.params.flat, enc.params.flat)


def test_reference_checkpoint_loads_after_conversion(M, tmp_path):
    """A checkpoint written by the reference's own save_checkpoint (under the mlx stand-in) -> converter -> load_checkpoint:
    the module mirrors then hold the reference's post-epoch weights and Adam moments."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle", "mlx_stub"))          # unpickling needs the array class
    try:
        spec = importlib.util.spec_from_file_location("conv", os.path.join(root, "tools", "convert_mlx_checkpoint.py"))
        conv = importlib.util.module_from_spec(spec); spec.loader.exec_module(conv)
        dst = str(tmp_path / "flat.npz")
        conv.main(os.path.join(root, "tests", "golden", "ref_checkpoint_tiny.npz"), dst)
    finally:
        sys.path.remove(os.path.join(root, "oracle", "mlx_stub"))
    ref = dict(np.load(os.path.join(root, "tests", "golden", "ref_tiny.npz")))
    kw = dict(vocab_size=11, embedding_dim=8, hidden_dim=16, latent_dim=8, num_conditions=1, num_layers=2)
    enc = M.MLXEncoder(**kw, seed=1); dec = M.MLXAutoregressiveDecoder(**kw, seed=2)
    tr = M.ARCVAETrainerWithLoss(enc, dec, None, None, learning_rate=1e-3, batch_size=5)
    assert tr.load_checkpoint(dst) == 3
    for k, v in ref.items():
        if k.startswith("penc/"):
            assert np.allclose(enc.params.views[k[5:]].cpu().numpy(), v, rtol=1e-6, atol=1e-7), k
        if k.startswith("pdec/"):
            assert np.allclose(dec.params.views[k[5:]].cpu().numpy(), v, rtol=1e-6, atol=1e-7), k
    k = "penc_opt/fc_mu.weight.v"
    view = enc.params.views["fc_mu.weight"]
    beg = (view.data_ptr() - enc.params.flat.data_ptr()) // 4
    got = tr.encoder_optimizer.v[beg:beg + view.numel()].reshape(view.shape).cpu().numpy()
    assert np.allclose(got, ref[k], rtol=1e-6, atol=1e-14)
    assert tr.history["epoch"] == [3.0]
