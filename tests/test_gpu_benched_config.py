"""Oracle parity AT THE BENCHED CONFIGURATION (BASELINE configs[1]: default dims, B=4096 x T=128), both precisions
(VERDICT r01 "What's weak" #3/#4).

  (a) per-row quantities (mu, logvar, logits) of 64 rows spread over the batch tiles of the B=4096 run against the
      fp64 oracle evaluated on exactly those rows;
  (b) the full step — 9 loss scalars and every gradient tensor of the GLOBAL batch — against ONE fp32 oracle step at
      B=4096 on the host cores (about a minute of CPU), teacher forcing everywhere so that bf16 near-ties of the greedy
      feedback cannot change the inputs (the feedback chain is checked separately in test_gpu_parity / test_gpu_bf16).

STATED TOLERANCES (relative to the largest magnitude of the tensor):
  fp32 mode: 1e-3 (north_star); measured ~1e-5
  bf16 mode: values 6e-3, gradients 2e-2  — twice what was measured at this configuration (3e-3 / 1e-2); the error
             budget covers bf16 operands of every tcgen05 contraction, bf16 tapes of the 128-step recurrence and
             MUFU.TANH activations (abs error 2^-11), i.e. more than the projections north_star names."""
import os

import numpy as np
import pytest
import torch

import arcvae_oracle as O
from _util import model_kwargs, rel_err

pytestmark = pytest.mark.gpu

B, T = 4096, 128
TOL = {"fp32": dict(val=1e-3, grad=1e-3), "bf16": dict(val=6e-3, grad=2e-2)}
HYPER = dict(beta=0.05, lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01, target_mi=4.85)
ROWS = np.r_[0:8, 120:136, 1017:1033, 2040:2056, 4088:4096]          # 64 rows, straddling 128-row tile borders


@pytest.fixture(scope="module")
def M():
    import mlx_vae_b200
    return mlx_vae_b200


@pytest.fixture(scope="module")
def setup():
    cfg = O.Config()
    p = O.init_params(cfg, seed=67, dtype=torch.float32)
    p["decoder"] = O.tree_map(lambda t: t * 3.0, p["decoder"])        # sharper logits than at initialisation
    x, cond, eps, _ = O.synthetic_batch(B, T, cfg, seed=67, tf_ratio=0.9)
    return cfg, p, x, cond, eps


def chunked_oracle_step(cfg, p, x, cond, eps, chunk=256):
    """The oracle's step on the whole batch with bounded memory.  Exact, not an approximation: the batch couples only
    through (mu, logvar) (losses/info.py:33-41), so (1) the encoder runs chunk by chunk without a tape, (2) the latent
    terms and their gradients w.r.t. the full [B,L] mu / logvar come from the oracle's loss functions, (3) every chunk is
    re-run with a tape and back-propagated with its rows of those cotangents plus its share of the CE sum."""
    NL, Bn, Tn = cfg.num_layers, x.shape[0], x.shape[1]
    xt, ct, et = torch.as_tensor(x), torch.as_tensor(cond), torch.as_tensor(eps)
    mask = np.ones(Tn, dtype=bool)
    with torch.no_grad():
        parts = [O.encoder_forward(p["encoder"], xt[i:i + chunk], ct[i:i + chunk], NL) for i in range(0, Bn, chunk)]
    mu = torch.cat([a for a, _ in parts]).requires_grad_(True)
    lv = torch.cat([b for _, b in parts]).requires_grad_(True)
    kl = O.kl_divergence(mu, lv, free_bits=HYPER["free_bits"])
    col = O.posterior_collapse(mu, lv, weight=HYPER["lambda_collapse"])
    mi = O.mutual_information(mu, lv)
    pen = HYPER["lambda_mi"] * O._mx_maximum(torch.zeros((), dtype=mi.dtype), HYPER["target_mi"] - mi)
    (HYPER["beta"] * kl + col + pen).backward()
    leaves = O.tree_map(lambda t: t.detach().clone().requires_grad_(True), p)
    ce_sum = 0.0
    for i in range(0, Bn, chunk):
        xc, cc = xt[i:i + chunk], ct[i:i + chunk]
        mu_c, lv_c = O.encoder_forward(leaves["encoder"], xc, cc, NL)
        lg = O.decoder_forward(leaves["decoder"], torch.zeros(xc.shape[0], cfg.latent_dim, dtype=ct.dtype), cc, NL, target_seq=xc, tf_mask=mask)
        ce = O.reconstruction_loss(lg, xc, reduction="sum")
        (ce / (Bn * Tn) + (mu_c * mu.grad[i:i + chunk]).sum() + (lv_c * lv.grad[i:i + chunk]).sum()).backward()
        ce_sum += float(ce.detach())
    recon = ce_sum / (Bn * Tn)
    vals = {"recon_loss": recon, "kl_loss": float(kl), "weighted_kl": HYPER["beta"] * float(kl), "collapse_penalty": float(col),
            "prop_loss": 0.0, "weighted_prop_loss": 0.0, "mutual_info": float(mi), "mi_penalty": float(pen),
            "total_loss": recon + HYPER["beta"] * float(kl) + float(col) + float(pen),
            "z": (mu + et * torch.exp(0.5 * lv)).detach()}
    grads = O.tree_map(lambda t: t.grad if t.grad is not None else torch.zeros_like(t), leaves)
    return vals, grads["encoder"], grads["decoder"]


@pytest.fixture(scope="module")
def oracle_step(setup):
    """ONE fp32 oracle step on the whole benched batch (all host threads, about a minute)."""
    cfg, p, x, cond, eps = setup
    torch.set_num_threads(os.cpu_count() or 1)
    return chunked_oracle_step(cfg, p, x, cond, eps)


def _modules(M, cfg, p, prec):
    kw = model_kwargs(cfg)
    enc = M.MLXEncoder(**kw, precision=prec).load_parameters(p["encoder"])
    dec = M.MLXAutoregressiveDecoder(**kw, precision=prec).load_parameters(p["decoder"])
    return enc, dec


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_rows_of_the_benched_batch_match_the_fp64_oracle(M, setup, prec):
    cfg, p, x, cond, eps = setup
    enc, dec = _modules(M, cfg, p, prec)
    dx, dc = torch.as_tensor(x).cuda(), torch.as_tensor(cond).cuda()
    mask = np.ones(T, dtype=bool)
    mu, logvar = enc(dx, dc)
    logits = dec(None, dc, target_seq=dx, tf_mask=mask)
    enc.check()
    p64 = O.tree_map(lambda t: t.double(), p)
    xs, cs = torch.as_tensor(x[ROWS]), torch.as_tensor(cond[ROWS]).double()
    mu_o, lv_o = O.encoder_forward(p64["encoder"], xs, cs, cfg.num_layers)
    lg_o = O.decoder_forward(p64["decoder"], torch.zeros(len(ROWS), cfg.latent_dim, dtype=torch.float64), cs, cfg.num_layers,
                             target_seq=xs, tf_mask=mask)
    idx = torch.as_tensor(ROWS).cuda()
    errs = dict(mu=rel_err(mu[idx].cpu(), mu_o), logvar=rel_err(logvar[idx].cpu(), lv_o), logits=rel_err(logits[idx].cpu(), lg_o))
    print(f"{prec} B={B} T={T}: 64-row slice vs fp64 oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < TOL[prec]["val"], (prec, k, v)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_full_step_at_the_benched_config_matches_the_fp32_oracle_step(M, setup, oracle_step, prec):
    cfg, p, x, cond, eps = setup
    vals, ge, gd = oracle_step
    enc, dec = _modules(M, cfg, p, prec)
    dx, dc, de = (torch.as_tensor(a).cuda() for a in (x, cond, eps))
    mask = np.ones(T, dtype=bool)
    d, (g_enc, g_dec) = M.loss_and_grad(enc, dec, None, dx, dc, eps=de, tf_mask=mask, **HYPER)
    enc.check()
    errs = {}
    for k in M._lib.LOSS_KEYS:
        errs[k] = abs(float(d[k]) - float(vals[k])) / max(abs(float(vals[k])), 1e-3)
        assert errs[k] < TOL[prec]["val"], (prec, k, float(d[k]), float(vals[k]))
    errs["z"] = rel_err(d["z"].cpu(), vals["z"])
    assert errs["z"] < TOL[prec]["val"]
    worst = ("", 0.0)
    for tree, ref, tag in ((g_enc, ge, "enc"), (g_dec, gd, "dec")):
        for mod, leaves in ref.items():
            for leaf, r in leaves.items():
                got = tree[mod][leaf].double().cpu()
                scale = float(r.abs().max())
                if scale == 0.0:
                    assert float(got.abs().max()) == 0.0, (mod, leaf)      # F1: structurally dead parameters
                    continue
                err = float((got - r.double()).abs().max()) / scale
                if err > worst[1]:
                    worst = (f"{tag}.{mod}.{leaf}", err)
                # the fp32 oracle itself carries ~1e-5 of summation noise over 524,288 positions
                assert err < TOL[prec]["grad"], (prec, tag, mod, leaf, err)
    print(f"{prec} B={B} T={T}: full step vs fp32 oracle step: worst scalar {max(errs.values()):.2e}, worst gradient {worst[0]} {worst[1]:.2e}")
