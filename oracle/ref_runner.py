"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference sources (/root/reference) under the ``mlx`` stand-in
(oracle/mlx_stub) and returns what they compute, as NumPy arrays, for a list of small seeded cases.

Used by ``oracle/make_ref_golden.py`` (writes tests/golden/ref_*.npz in THIS container, where /root/reference is mounted)
and by ``tests/test_ref_pin.py`` (re-runs it live when the mount is present).  Nothing on the product path imports this.

Each case exercises, through the reference's own code:
  models/encoder.py:76-132, :134-155     MLXEncoder.__call__, reparameterize
  models/decoder.py:76-111, :113-190     initialize_hidden_state, MLXAutoregressiveDecoder.__call__ (host coins, :180)
  models/decoder_sampling.py:48-128      generate_with_temperature (early stopping on / off)
  models/vae.py:63-99                    ARCVAE.__call__
  losses/{recon,kl,info,prop}.py         every loss function, also at the shapes of test_loss_signs.py:19-23
  complete_vae_loss.py:7-99              the 12-key dict
  trainer.py:242-416, :489-522           _train_epoch_batches (value_and_grad, no-op clip, 2x Adam, logging pass), _clip_gradients
  mlx_data/dataloader.py:4-111           MoleculeDataset normalisation, padding, shuffled ragged batches
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
STUB = os.path.join(HERE, "mlx_stub")

if HERE not in sys.path:
    sys.path.insert(0, HERE)
import arcvae_oracle as O  # noqa: E402

CASES = {
    # name: (Config kwargs, B, T, tf_ratio, decoder weight scale, seed)
    "tiny": (dict(vocab_size=11, embedding_dim=8, hidden_dim=16, latent_dim=8, num_conditions=1, num_layers=2), 5, 7, 0.6, 4.0, 101),
    "tiny_c2l3": (dict(vocab_size=13, embedding_dim=8, hidden_dim=12, latent_dim=6, num_conditions=2, num_layers=3), 4, 6, 0.5, 5.0, 102),
    "default_b4": (dict(vocab_size=80, embedding_dim=128, hidden_dim=256, latent_dim=128, num_conditions=1, num_layers=2), 4, 10, 0.9, 3.0, 103),
}
HYPER = dict(beta=0.05, lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01, target_mi=4.85)
TRAINER_HYPER = dict(lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01)
LR = 2e-3


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "models"))


@contextlib.contextmanager
def reference_imports():
    """sys.path with the stub ``mlx`` and the reference tree in front; yields the imported reference modules."""
    added = [STUB, REFERENCE]
    for p in reversed(added):
        sys.path.insert(0, p)
    try:
        import mlx.core as mx
        import mlx.nn  # noqa: F401
        mx.set_default_float(torch.float64)
        import complete_vae_loss as ref_cvl
        import losses as ref_losses
        import mlx_data.dataloader as ref_data
        import models as ref_models
        import trainer as ref_trainer
        assert ref_models.__file__.startswith(REFERENCE) and ref_trainer.__file__.startswith(REFERENCE)
        yield dict(mx=mx, models=ref_models, losses=ref_losses, cvl=ref_cvl, trainer=ref_trainer, data=ref_data)
    finally:
        for p in added:
            sys.path.remove(p)


def case_inputs(name):
    """Seeded inputs shared by the reference run and the oracle run (fp64)."""
    kw, B, T, tf, dscale, seed = CASES[name]
    cfg = O.Config(**kw)
    params = O.init_params(cfg, seed=seed, dtype=torch.float64)
    params["decoder"] = O.tree_map(lambda t: t * dscale, params["decoder"])
    x, cond, eps, _ = O.synthetic_batch(B, T, cfg, seed=seed + 1, tf_ratio=tf)
    rng = np.random.default_rng(seed + 2)
    n_mol = 2 * B + 3                                                   # 3 batches, the last one ragged
    lens = rng.integers(2, T + 3, size=n_mol)                           # some longer than T: truncation path
    mols = [[int(v) for v in rng.integers(3, cfg.vocab_size, size=int(n) - 1)] + [cfg.end_token] for n in lens]
    props = rng.normal(60.0, 25.0, size=(n_mol, cfg.num_conditions)).astype(np.float32)
    eps_q = rng.standard_normal((4, B, cfg.latent_dim))                 # draws consumed by the trainer epoch, in order
    return dict(cfg=cfg, B=B, T=T, tf=tf, seed=seed, params=params, x=x, cond=cond.astype(np.float64),
                eps=eps.astype(np.float64), mols=mols, props=props, eps_q=eps_q)


def _load_into(mx, module, tree):
    """Same mechanism as the reference's ``_load_module_weights`` (trainer.py:714-726): setattr of mx.arrays."""
    for k, v in tree.items():
        if isinstance(v, dict):
            _load_into(mx, getattr(module, k), v)
        else:
            setattr(module, k, mx.array(v.detach().clone().numpy()))


def _tree_np(tree):
    return {k: (_tree_np(v) if isinstance(v, dict) else np.asarray(v.detach().cpu().numpy())) for k, v in tree.items()
            if isinstance(v, (dict, torch.Tensor))}


def _flat(prefix, tree, out):
    for k, v in tree.items():
        if isinstance(v, dict):
            _flat(f"{prefix}{k}.", v, out)
        else:
            out[f"{prefix}{k}"] = v
    return out


def summarize(a: np.ndarray, seed: int = 0):
    """Compact signature of a large tensor: [sum, sum of squares, sum of |x|*index weights] + 64 sampled elements."""
    f = np.asarray(a, dtype=np.float64).reshape(-1)
    idx = np.random.default_rng(seed).integers(0, f.size, size=min(64, f.size))
    w = np.cos(np.arange(f.size) * 0.7)
    return np.concatenate([[f.sum(), (f * f).sum(), (f * w).sum()], f[idx]])


def run_reference(name, checkpoint_to=None):
    """Everything the reference computes for case `name`, keyed like the fixture file."""
    ci = case_inputs(name)
    cfg, B, T, tf, seed = ci["cfg"], ci["B"], ci["T"], ci["tf"], ci["seed"]
    out = {}
    with reference_imports() as R:
        mx = R["mx"]
        kw = dict(vocab_size=cfg.vocab_size, embedding_dim=cfg.embedding_dim, hidden_dim=cfg.hidden_dim,
                  latent_dim=cfg.latent_dim, num_conditions=cfg.num_conditions, num_layers=cfg.num_layers)
        vae = R["models"].ARCVAE(**kw)
        _load_into(mx, vae.encoder, ci["params"]["encoder"])
        _load_into(mx, vae.decoder, ci["params"]["decoder"])
        _load_into(mx, vae.decoder_sampling.decoder, ci["params"]["decoder"])
        x = mx.array(ci["x"], dtype=mx.uint32)
        cond = mx.array(ci["cond"])
        eps = mx.array(ci["eps"])

        mu, logvar = vae.encoder(x, cond)                                           # encoder.py:76
        out["mu"], out["logvar"] = mu.numpy(), logvar.numpy()
        mx.random.inject([eps])
        out["z"] = R["models"].MLXEncoder.reparameterize(mu, logvar).numpy()         # encoder.py:134
        hid, cell = vae.decoder.initialize_hidden_state(mx.array(out["z"]), cond)    # decoder.py:76
        out["hidden_init"], out["cell_init_absmax"] = hid.numpy(), np.abs(cell.numpy()).max()

        np.random.seed(seed)                                                          # coins of decoder.py:180
        logits = vae.decoder(mx.array(out["z"]), cond, target_seq=x, teacher_forcing_ratio=tf)
        out["logits"] = logits.numpy()
        np.random.seed(seed)
        out["logits_tf0"] = vae.decoder(mx.array(out["z"]), cond, target_seq=x, teacher_forcing_ratio=0.0).numpy()
        out["logits_free"] = vae.decoder(mx.array(out["z"]), cond, target_seq=None, max_length=T + 2).numpy()

        np.random.seed(seed)
        mx.random.inject([eps])
        lg, m2, lv2, z2 = vae(x, cond, target_seq=x, teacher_forcing_ratio=tf)      # vae.py:63
        out["vae_logits"], out["vae_z"] = lg.numpy(), z2.numpy()

        samp = vae.decoder_sampling
        out["tokens_early"] = samp.generate_with_temperature(mx.array(out["z"]), cond, max_length=T + 5, temperature=0.7).numpy()
        out["tokens_full"] = samp.generate_with_temperature(mx.array(out["z"]), cond, max_length=T + 5, temperature=1.3,
                                                            early_stopping=False).numpy()

        # early stopping (decoder_sampling.py:87-88): with a dominant end-token bias every row ends at step 0
        b_out = samp.decoder.fc_out.bias
        samp.decoder.fc_out.bias = b_out + mx.array(np.eye(cfg.vocab_size)[cfg.end_token] * 50.0)
        out["tokens_stop"] = samp.generate_with_temperature(mx.array(out["z"]), cond, max_length=T + 5, temperature=1.0).numpy()
        samp.decoder.fc_out.bias = b_out

        np.random.seed(seed)
        mx.random.inject([eps])
        d = R["cvl"].complete_vae_loss(vae.encoder, vae.decoder, None, x, cond, teacher_forcing_ratio=tf, **HYPER)
        for k, v in d.items():
            out["loss/" + k] = v.numpy()

        # gradients exactly as trainer.py:292-305 takes them
        def model_loss_fn(encoder, decoder, xx, cc):
            return R["cvl"].complete_vae_loss(encoder, decoder, None, xx, cc, teacher_forcing_ratio=tf, **HYPER)["total_loss"]
        np.random.seed(seed)
        mx.random.inject([eps])
        loss, grads = mx.value_and_grad(model_loss_fn, argnums=[0, 1])(vae.encoder, vae.decoder, x, cond)
        out["vg_loss"] = loss.numpy()
        assert isinstance(grads, tuple) and type(grads[0]) is dict and type(grads[1]["fc_out"]) is dict
        genc, gdec = _flat("", _tree_np(grads[0]), {}), _flat("", _tree_np(grads[1]), {})
        big = cfg.hidden_dim >= 128
        for tag, g in (("genc", genc), ("gdec", gdec)):
            for k, v in g.items():
                out[f"{tag}/{k}"] = summarize(v) if big else v
        # trainer.py:489-522: the clip sees no top-level arrays -> norm 0 -> gradients returned untouched (F4)
        scaled = tuple({k: {kk: vv * 1e6 for kk, vv in v.items()} for k, v in g.items()} for g in grads)
        clipped = R["trainer"].ARCVAETrainerWithLoss._clip_gradients(scaled, 1.0)
        out["clip_is_noop"] = np.array(all(torch.equal(clipped[i][k][kk], scaled[i][k][kk])
                                           for i in range(2) for k in scaled[i] for kk in scaled[i][k]))

        # losses at the shapes of test_loss_signs.py:19-23 (seeded here; the script itself is unseeded)
        rs = np.random.default_rng(seed + 7)
        ls_logits = rs.standard_normal((32, 120, 95)); ls_t = rs.integers(0, 95, (32, 120))
        ls_mu = rs.standard_normal((32, 128)) * 0.1; ls_lv = rs.standard_normal((32, 128)) * 0.1 - 1.0
        L = R["losses"]
        out["signs/recon"] = L.reconstruction_loss(mx.array(ls_logits), mx.array(ls_t, dtype=mx.uint32), reduction="mean").numpy()
        out["signs/recon_sum"] = L.reconstruction_loss(mx.array(ls_logits), mx.array(ls_t, dtype=mx.uint32), reduction="sum").numpy()
        out["signs/kl_fb0"] = L.kl_divergence(mx.array(ls_mu), mx.array(ls_lv), reduction="mean", free_bits=0.0).numpy()
        out["signs/kl_fb1"] = L.kl_divergence(mx.array(ls_mu), mx.array(ls_lv), reduction="mean", free_bits=1.0).numpy()
        out["signs/kl_none"] = L.kl_divergence(mx.array(ls_mu), mx.array(ls_lv), reduction="none", free_bits=0.5).numpy()
        out["signs/mi"] = L.mutual_information(mx.array(ls_mu), mx.array(ls_lv)).numpy()
        out["signs/collapse"] = L.posterior_collapse(mx.array(ls_mu), mx.array(ls_lv), target_mi=4.85, weight=0.1).numpy()
        pp, tp = rs.standard_normal((32, 1)), rs.standard_normal((32, 1))
        out["signs/prop"] = L.property_prediction_loss(None, mx.array(pp), mx.array(tp), reduction="mean").numpy()
        out["signs/prop_scaled"] = L.property_prediction_loss(None, mx.array(pp), mx.array(tp), mx.array([[2.0]]), reduction="sum").numpy()

        # dataset + one epoch of the reference trainer (3 batches, the last ragged; logging pass after batch 0)
        ds = R["data"].MoleculeDataset(ci["mols"], ci["props"], max_length=T, pad_token=0)
        out["ds/props_norm"] = np.asarray(ds.properties_normalized)
        out["ds/item3_mol"] = ds[3]["molecule"].numpy()
        np.random.seed(seed + 3)
        batches = list(ds.to_batches(B, shuffle=True))
        out["ds/n_batches"] = np.array(len(batches))
        out["ds/batch0_mol"], out["ds/batch0_prop"] = batches[0][0].numpy(), batches[0][1].numpy()
        out["ds/last_mol"] = batches[-1][0].numpy()
        with tempfile.TemporaryDirectory() as tmp:
            tr = R["trainer"].ARCVAETrainerWithLoss(vae.encoder, vae.decoder, None, ds, learning_rate=LR, batch_size=B,
                                                    checkpoint_dir=os.path.join(tmp, "ck"), **TRAINER_HYPER)
            out["beta_e3_of_100"] = np.array(tr.compute_beta(3))
            out["tf_e3_of_30"] = np.array(tr.compute_teacher_forcing_ratio(3, 30))
            np.random.seed(seed + 4)
            q = [mx.array(e) for e in ci["eps_q"]]
            q[3] = mx.array(ci["eps_q"][3][: len(ds) - 2 * B])                      # ragged last batch
            mx.random.inject(q)
            ep = tr._train_epoch_batches(beta=HYPER["beta"], teacher_forcing_ratio=tf)   # trainer.py:242
            assert mx.random.pending() == 0
            if checkpoint_to is not None:
                # trainer.py:577-603: np.savez of nested dicts of mx.array (pickled object arrays)
                tr.history["epoch"].append(3); tr.history["train_loss"].append(float(ep["loss"]))
                tr.save_checkpoint(3)
                import shutil
                shutil.copyfile(os.path.join(tmp, "ck", "checkpoint_epoch_003.npz"), checkpoint_to)
        for k, v in ep.items():
            out["epoch/" + k] = np.array(v)
        for tag, mod, opt in (("penc", vae.encoder, tr.encoder_optimizer), ("pdec", vae.decoder, tr.decoder_optimizer)):
            for k, v in _flat("", _tree_np(mod.parameters()), {}).items():
                out[f"{tag}/{k}"] = summarize(v) if big else v
            st = {k: v for k, v in opt.state.items() if isinstance(v, dict)}
            for k, v in _flat("", _tree_np(st), {}).items():
                out[f"{tag}_opt/{k}"] = summarize(v) if big else v
    return {k: np.asarray(v) for k, v in out.items()}


def run_oracle(name):
    """The same quantities from oracle/arcvae_oracle.py + oracle/dataset_oracle.py (fp64)."""
    import dataset_oracle as D
    ci = case_inputs(name)
    cfg, B, T, tf, seed, p = ci["cfg"], ci["B"], ci["T"], ci["tf"], ci["seed"], ci["params"]
    NL = cfg.num_layers
    x = torch.as_tensor(ci["x"]).long()
    cond, eps = torch.as_tensor(ci["cond"]), torch.as_tensor(ci["eps"])
    out = {}
    mu, logvar = O.encoder_forward(p["encoder"], x, cond, NL)
    out["mu"], out["logvar"] = mu.numpy(), logvar.numpy()
    z = O.reparameterize(mu, logvar, eps)
    out["z"] = z.numpy()
    hid, cell = O.initialize_hidden_state(p["decoder"], z, cond, NL)
    out["hidden_init"], out["cell_init_absmax"] = hid.numpy(), np.abs(cell.numpy()).max()

    def coins(s, n, ratio):
        np.random.seed(s)
        return np.array([np.random.rand() < ratio for _ in range(n)])
    mask = coins(seed, T, tf)
    out["logits"] = O.decoder_forward(p["decoder"], z, cond, NL, target_seq=x, tf_mask=mask).numpy()
    out["logits_tf0"] = O.decoder_forward(p["decoder"], z, cond, NL, target_seq=x, tf_mask=np.zeros(T, bool)).numpy()
    out["logits_free"] = O.decoder_forward(p["decoder"], z, cond, NL, target_seq=None, max_length=T + 2).numpy()
    lg, _, _, z2 = O.vae_forward(p, x, cond, NL, eps, target_seq=x, tf_mask=mask)
    out["vae_logits"], out["vae_z"] = lg.numpy(), z2.numpy()
    out["tokens_early"] = O.generate_with_temperature(p["decoder"], z, cond, NL, max_length=T + 5, temperature=0.7).numpy()
    out["tokens_full"] = O.generate_with_temperature(p["decoder"], z, cond, NL, max_length=T + 5, temperature=1.3,
                                                     early_stopping=False).numpy()
    p_stop = {k: dict(v) for k, v in p["decoder"].items()}
    p_stop["fc_out"]["bias"] = p_stop["fc_out"]["bias"] + torch.as_tensor(np.eye(cfg.vocab_size)[cfg.end_token] * 50.0)
    out["tokens_stop"] = O.generate_with_temperature(p_stop, z, cond, NL, max_length=T + 5, temperature=1.0).numpy()
    d = O.complete_vae_loss(p, x, cond, NL, eps, mask, **HYPER)
    for k, v in d.items():
        out["loss/" + k] = v.numpy()
    vals, (ge, gd) = O.loss_and_grads(p, x, cond, NL, eps, mask, **HYPER)
    out["vg_loss"] = vals["total_loss"].numpy()
    big = cfg.hidden_dim >= 128
    for tag, g in (("genc", ge), ("gdec", gd)):
        for k, v in O.tree_flatten(g).items():
            out[f"{tag}/{k}"] = summarize(v.numpy()) if big else v.numpy()
    scaled = tuple(O.tree_map(lambda t: t * 1e6, g) for g in (ge, gd))
    clipped = O.clip_gradients_reference(scaled, 1.0)
    out["clip_is_noop"] = np.array(all(torch.equal(a, b) for i in range(2)
                                       for a, b in zip(O.tree_flatten(clipped[i]).values(), O.tree_flatten(scaled[i]).values())))
    rs = np.random.default_rng(seed + 7)
    ls_logits = torch.as_tensor(rs.standard_normal((32, 120, 95))); ls_t = torch.as_tensor(rs.integers(0, 95, (32, 120)))
    ls_mu = torch.as_tensor(rs.standard_normal((32, 128)) * 0.1); ls_lv = torch.as_tensor(rs.standard_normal((32, 128)) * 0.1 - 1.0)
    out["signs/recon"] = O.reconstruction_loss(ls_logits, ls_t).numpy()
    out["signs/recon_sum"] = O.reconstruction_loss(ls_logits, ls_t, reduction="sum").numpy()
    out["signs/kl_fb0"] = O.kl_divergence(ls_mu, ls_lv, free_bits=0.0).numpy()
    out["signs/kl_fb1"] = O.kl_divergence(ls_mu, ls_lv, free_bits=1.0).numpy()
    out["signs/kl_none"] = O.kl_divergence(ls_mu, ls_lv, reduction="none", free_bits=0.5).numpy()
    out["signs/mi"] = O.mutual_information(ls_mu, ls_lv).numpy()
    out["signs/collapse"] = O.posterior_collapse(ls_mu, ls_lv, target_mi=4.85, weight=0.1).numpy()
    pp, tp = torch.as_tensor(rs.standard_normal((32, 1))), torch.as_tensor(rs.standard_normal((32, 1)))
    out["signs/prop"] = O.property_prediction_loss(None, pp, tp).numpy()
    out["signs/prop_scaled"] = O.property_prediction_loss(None, pp, tp, torch.tensor([[2.0]], dtype=torch.float64), reduction="sum").numpy()

    ds = D.MoleculeDatasetOracle(ci["mols"], ci["props"], max_length=T, pad_token=0)
    out["ds/props_norm"] = np.asarray(ds.properties_normalized)
    out["ds/item3_mol"] = ds[3]["molecule"].astype(np.int64)
    np.random.seed(seed + 3)
    batches = list(ds.to_batches(B, shuffle=True))
    out["ds/n_batches"] = np.array(len(batches))
    out["ds/batch0_mol"], out["ds/batch0_prop"] = batches[0][0].astype(np.int64), batches[0][1]
    out["ds/last_mol"] = batches[-1][0].astype(np.int64)
    out["beta_e3_of_100"] = np.array(0.0 + (0.4 - 0.0) * (3 / 100))                 # trainer.py:102-108 with its defaults
    out["tf_e3_of_30"] = np.array(max(0.5, 0.9 - 0.4 * (3 / 30)))                   # trainer.py:110-114

    # the epoch of trainer.py:242-416: shuffle, then per batch T coins + value_and_grad + (no-op) clip + 2 x Adam;
    # after batch 0 (and every 25th) a second forward for logging that draws T more coins and another eps
    np.random.seed(seed + 4)
    params, state = p, O.adam_init(p)
    hyper_tr = dict(HYPER)
    total, nb, q = 0.0, 0, 0
    comp = None
    for bi, (mb, pb) in enumerate(ds.to_batches(B, shuffle=True)):
        xb = torch.as_tensor(mb.astype(np.int64)); cb = torch.as_tensor(pb).double()
        m = np.array([np.random.rand() < tf for _ in range(T)])
        e = torch.as_tensor(ci["eps_q"][q][: xb.shape[0]]); q += 1
        vals, grads, params, state = O.train_step(params, state, xb, cb, NL, e, m, LR, **hyper_tr)
        if bi == 0 or bi % 25 == 0:                                                  # trainer.py:336-363
            m2 = np.array([np.random.rand() < tf for _ in range(T)])
            e2 = torch.as_tensor(ci["eps_q"][q][: xb.shape[0]]); q += 1
            dd = O.complete_vae_loss(params, xb, cb, NL, e2, m2, **hyper_tr)
            comp = {k: float(dd[k]) for k in ("recon_loss", "kl_loss", "collapse_penalty", "prop_loss")}
        total += float(vals["total_loss"]); nb += 1
    out["epoch/loss"] = np.array(total / nb)
    out["epoch/recon"], out["epoch/kl"] = np.array(comp["recon_loss"]), np.array(comp["kl_loss"])
    out["epoch/collapse"], out["epoch/prop"] = np.array(comp["collapse_penalty"]), np.array(comp["prop_loss"])
    for tag, key in (("penc", "encoder"), ("pdec", "decoder")):
        for k, v in O.tree_flatten(params[key]).items():
            out[f"{tag}/{k}"] = summarize(v.numpy()) if big else v.numpy()
        for mom in ("m", "v"):
            for k, v in O.tree_flatten(state[mom][key]).items():
                out[f"{tag}_opt/{k}.{mom}"] = summarize(v.numpy()) if big else v.numpy()
    return {k: np.asarray(v) for k, v in out.items()}


def compare(a: dict, b: dict, rtol=1e-10):
    """List of (key, error) where the two result dicts disagree; integer / bool entries must be identical."""
    bad = []
    for k in sorted(set(a) | set(b)):
        if k not in a or k not in b:
            bad.append((k, "missing"))
            continue
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape:
            bad.append((k, f"shape {x.shape} vs {y.shape}"))
        elif x.dtype.kind in "iub" or y.dtype.kind in "iub":
            if not np.array_equal(x.astype(np.int64), y.astype(np.int64)):
                bad.append((k, "integer mismatch"))
        else:
            err = float(np.abs(x - y).max()) if x.size else 0.0
            den = max(float(np.abs(y).max()) if y.size else 0.0, 1e-300)
            if not err <= rtol * max(den, 1e-6):
                bad.append((k, err / den))
    return bad
