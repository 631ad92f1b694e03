"""GPU tests of the persistent cluster LSTM recurrence (csrc/lstm_cluster.cu, hidden_dim 256, precision='bf16')
against the per-step tensor-core path and the fp32 path of the same library, and against the fp64 fixture."""
import os

import numpy as np
import pytest
import torch

import arcvae_oracle as O
from _util import golden_cfg, golden_hyper, golden_params, load_golden, model_kwargs, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mlx_vae_b200
    return mlx_vae_b200


def cuda(a):
    return torch.as_tensor(np.asarray(a)).cuda()


def run_encoder(M, p, x, cond, prec, cluster, dmu=None, dlv=None):
    if cluster:
        os.environ.pop("ARCVAE_NO_CLUSTER", None)
    else:
        os.environ["ARCVAE_NO_CLUSTER"] = "1"
    try:
        enc = M.MLXEncoder(**model_kwargs(O.Config()), precision=prec).load_parameters(p["encoder"])
        mu, lv = enc(cuda(x), cuda(cond))
        enc.check()
        grads = None
        if dmu is not None:
            enc.zero_grad()
            enc.backward(dmu, dlv)
            enc.check()
            grads = {k: v.clone() for k, v in O.tree_flatten(enc.gradients()).items()}
        torch.cuda.synchronize()
        return mu.clone(), lv.clone(), grads
    finally:
        os.environ.pop("ARCVAE_NO_CLUSTER", None)


@pytest.mark.parametrize("B,T", [(64, 1), (130, 3), (128, 2), (128, 9), (256, 24), (200, 16), (4096, 128)])
def test_cluster_matches_per_step_paths(M, B, T):
    cfg = O.Config()
    p = O.init_params(cfg, seed=3, dtype=torch.float32)
    x, cond, eps, _ = O.synthetic_batch(B, T, cfg, seed=B + T)
    g = torch.Generator(device="cuda").manual_seed(1)
    dmu = torch.randn(B, cfg.latent_dim, device="cuda", generator=g) / B
    dlv = torch.randn(B, cfg.latent_dim, device="cuda", generator=g) / B
    mu32, lv32, g32 = run_encoder(M, p, x, cond, "fp32", False, dmu, dlv)
    muS, lvS, gS = run_encoder(M, p, x, cond, "bf16", False, dmu, dlv)
    muC, lvC, gC = run_encoder(M, p, x, cond, "bf16", True, dmu, dlv)
    e_fwd = max(rel_err(muC.cpu(), mu32.cpu()), rel_err(lvC.cpu(), lv32.cpu()))
    e_fwd_step = max(rel_err(muS.cpu(), mu32.cpu()), rel_err(lvS.cpu(), lv32.cpu()))
    assert e_fwd < 8e-3, (e_fwd, e_fwd_step)
    worst, worst_step, name = 0.0, 0.0, None
    for n in g32:
        s = float(g32[n].abs().max())
        if s == 0.0:
            assert float(gC[n].abs().max()) == 0.0
            continue
        e = float((gC[n] - g32[n]).abs().max()) / s
        if e > worst:
            worst, name = e, n
        worst_step = max(worst_step, float((gS[n] - g32[n]).abs().max()) / s)
    print(f"B={B} T={T}: fwd err cluster {e_fwd:.2e} / per-step {e_fwd_step:.2e}; grad err cluster {worst:.2e} ({name}) "
          f"/ per-step {worst_step:.2e}")
    assert worst < 2e-2, (worst, name)


def test_cluster_path_against_fp64_fixture(M):
    g = load_golden("default_b8")
    cfg = golden_cfg(g); p = golden_params(g); kw = model_kwargs(cfg); hyper = golden_hyper(g)
    enc = M.MLXEncoder(**kw, precision="bf16").load_parameters(p["encoder"])
    mu, lv = enc(cuda(g["x"]), cuda(g["cond"]))
    enc.check()
    assert rel_err(mu.cpu(), g["mu"]) < 8e-3 and rel_err(lv.cpu(), g["logvar"]) < 8e-3
