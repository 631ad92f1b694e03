"""carry_state=True decoder — an extension, not the reference's behaviour: the state built by
initialize_hidden_state is consumed and carried across positions.  Oracle: oracle/arcvae_oracle.py with
carry_state=True (itself cross-checked against torch.nn.LSTM with an initial state, tests/test_oracle.py).

Checked through the C ABI: logits and fed tokens (teacher forcing, coins, pure feedback), the full loss dict and every
gradient — including z_to_hidden / condition_to_hidden / Wh, which are exactly zero in the reference mode, and the
encoder gradients that now receive the reconstruction loss through z — and the greedy sampler with state."""
import numpy as np
import pytest
import torch

import arcvae_oracle as O
from _util import model_kwargs, rel_err

pytestmark = pytest.mark.gpu
HYPER = dict(beta=0.05, lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01, target_mi=4.85)


@pytest.fixture(scope="module")
def M():
    import mlx_vae_b200
    return mlx_vae_b200


def cuda(a):
    return torch.as_tensor(np.asarray(a)).cuda()


@pytest.mark.parametrize("dims,B,T,tf", [((13, 8, 12, 6, 2, 3), 5, 7, 0.6), ((80, 128, 256, 128, 1, 2), 16, 9, 1.0),
                                        ((23, 16, 64, 16, 1, 2), 33, 6, 0.0)])
def test_carry_state_step_matches_oracle_fp32(M, dims, B, T, tf):
    cfg = O.Config(*dims)
    p = O.init_params(cfg, seed=31, dtype=torch.float64)
    p["decoder"] = O.tree_map(lambda t: t * 2.0, p["decoder"])
    x, cond, eps, mask = O.synthetic_batch(B, T, cfg, seed=32, tf_ratio=tf)
    if tf == 1.0:
        mask = np.ones(T, dtype=bool)
    xt, ct, et = torch.as_tensor(x), torch.as_tensor(cond).double(), torch.as_tensor(eps).double()
    vals, (ge, gd) = O.loss_and_grads(p, xt, ct, cfg.num_layers, et, mask, carry_state=True, **HYPER)
    lg_o, in_o = O.decoder_forward(p["decoder"], vals["z"], ct, cfg.num_layers, target_seq=xt, tf_mask=mask,
                                   return_inputs=True, carry_state=True)
    kw = model_kwargs(cfg)
    enc = M.MLXEncoder(**kw).load_parameters(p["encoder"])
    dec = M.MLXAutoregressiveDecoder(**kw, carry_state=True).load_parameters(p["decoder"])
    d, (g_enc, g_dec) = M.loss_and_grad(enc, dec, None, cuda(x), cuda(cond), eps=cuda(eps), tf_mask=mask, **HYPER)
    assert torch.equal(dec.last_inputs.cpu().long(), in_o), "tokens fed to the decoder differ from the oracle"
    lg = dec(d["z"], cuda(cond), target_seq=cuda(x), tf_mask=mask)
    assert rel_err(lg.cpu(), lg_o) < 1e-3
    for k in M._lib.LOSS_KEYS:
        assert abs(float(d[k]) - float(vals[k])) <= 1e-4 * max(1.0, abs(float(vals[k]))), (k, float(d[k]), float(vals[k]))
    worst = ("", 0.0)
    for tree, ref, tag in ((g_enc, ge, "enc"), (g_dec, gd, "dec")):
        for mod, leaves in ref.items():
            for leaf, r in leaves.items():
                got = tree[mod][leaf].double().cpu()
                s = float(r.abs().max())
                if s == 0.0:
                    assert float(got.abs().max()) == 0.0, (mod, leaf)
                    continue
                e = float((got - r).abs().max()) / s
                if e > worst[1]:
                    worst = (f"{tag}.{mod}.{leaf}", e)
                assert e < 1e-3, (tag, mod, leaf, e)
    # the parameters that are dead in the reference mode are alive here
    for mod, leaf in (("z_to_hidden", "weight"), ("condition_to_hidden", "bias"), ("lstm_layer_0", "Wh")):
        assert float(g_dec[mod][leaf].abs().max()) > 0, (mod, leaf)
    print(f"carry_state dims={dims} B={B} T={T} tf={tf}: worst gradient {worst[0]} {worst[1]:.2e}")


def test_carry_state_bf16_within_tolerance(M):
    cfg = O.Config()
    B, T = 128, 8
    p = O.init_params(cfg, seed=33, dtype=torch.float64)
    x, cond, eps, _ = O.synthetic_batch(B, T, cfg, seed=34)
    mask = np.ones(T, dtype=bool)
    xt, ct, et = torch.as_tensor(x), torch.as_tensor(cond).double(), torch.as_tensor(eps).double()
    vals, (ge, gd) = O.loss_and_grads(p, xt, ct, cfg.num_layers, et, mask, carry_state=True, **HYPER)
    kw = model_kwargs(cfg)
    enc = M.MLXEncoder(**kw, precision="bf16").load_parameters(p["encoder"])
    dec = M.MLXAutoregressiveDecoder(**kw, precision="bf16", carry_state=True).load_parameters(p["decoder"])
    d, (g_enc, g_dec) = M.loss_and_grad(enc, dec, None, cuda(x), cuda(cond), eps=cuda(eps), tf_mask=mask, **HYPER)
    for k in ("total_loss", "recon_loss", "kl_loss"):
        assert abs(float(d[k]) - float(vals[k])) <= 8e-3 * max(1.0, abs(float(vals[k]))), k
    for tree, ref in ((g_enc, ge), (g_dec, gd)):
        for mod, leaves in ref.items():
            for leaf, r in leaves.items():
                s = float(r.abs().max())
                if s > 0:
                    assert float((tree[mod][leaf].double().cpu() - r).abs().max()) / s < 2e-2, (mod, leaf)


def test_carry_state_sampler_and_vae(M):
    cfg = O.Config(23, 16, 64, 16, 1, 2)
    p = O.init_params(cfg, seed=35, dtype=torch.float64)
    p["decoder"] = O.tree_map(lambda t: t * 3.0, p["decoder"])
    B, T = 40, 10
    g = torch.Generator().manual_seed(1)
    z = torch.randn(B, cfg.latent_dim, generator=g, dtype=torch.float64)
    cond = torch.randn(B, 1, generator=g, dtype=torch.float64)
    ref = O.generate_with_temperature(p["decoder"], z, cond, cfg.num_layers, max_length=T, early_stopping=False, carry_state=True)
    ref_es = O.generate_with_temperature(p["decoder"], z, cond, cfg.num_layers, max_length=T, carry_state=True)
    kw = model_kwargs(cfg)
    s = M.MLXAutoregressiveDecoderSampling(**kw, carry_state=True)
    s.decoder.load_parameters(p["decoder"])
    out = s.generate_with_temperature(z.float().cuda(), cond.float().cuda(), max_length=T, early_stopping=False)
    assert out.dtype == torch.int32 and torch.equal(out.cpu().long(), ref)
    out_es = s.generate_with_temperature(z.float().cuda(), cond.float().cuda(), max_length=T)
    assert torch.equal(out_es.cpu().long(), ref_es)
    with pytest.raises(NotImplementedError):
        s.generate_with_temperature(z.float().cuda(), cond.float().cuda(), max_length=T, multinomial=True)
    with pytest.raises(ValueError):
        s.decoder(None, cond.float().cuda(), max_length=3)
    vae = M.ARCVAE(**kw, seed=2, carry_state=True)
    toks = vae.generate(B, cond.float().cuda(), max_length=6, seed=5)
    toks2 = vae.generate(B, cond.float().cuda(), max_length=6, seed=6)
    assert toks.shape[0] == B and toks.shape[1] <= 6
    lg, mu, lv, zz = vae(torch.randint(0, 23, (B, 5), device="cuda"), cond.float().cuda(),
                         target_seq=torch.randint(0, 23, (B, 5), device="cuda"), tf_mask=np.ones(5, dtype=bool), seed=3)
    assert tuple(lg.shape) == (B, 5, 23) and torch.isfinite(lg).all()
    assert toks2.shape[0] == B
