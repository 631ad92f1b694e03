"""CPU tests of the oracle itself (oracle/arcvae_oracle.py).

The reference pins no numbers for this path (no tests, MLX not installable: parity unpinned), so the oracle is
validated by independent means: the committed fp64 fixtures, torch.nn.LSTM, finite differences, the invariants the
reference implies and a Philox known-answer vector."""
import numpy as np
import pytest
import torch

import arcvae_oracle as O
import philox_ref
from _util import golden_cfg, golden_hyper, golden_params, load_golden, rel_err

FULL = ["tiny", "tiny_c2l3", "tiny_sharp"]
SAMPLED = ["default_b8", "default_b8_sharp"]


def _run(g, dtype):
    cfg = golden_cfg(g)
    p = golden_params(g, dtype)
    x = torch.as_tensor(g["x"])
    c = torch.as_tensor(g["cond"]).to(dtype)
    e = torch.as_tensor(g["eps"]).to(dtype)
    return cfg, p, O.loss_and_grads(p, x, c, cfg.num_layers, e, g["tf_mask"], **golden_hyper(g))


@pytest.mark.parametrize("name", FULL + SAMPLED)
def test_fp64_oracle_reproduces_fixture(name):
    g = load_golden(name)
    cfg, p, (vals, (ge, gd)) = _run(g, torch.float64)
    assert rel_err(vals["logits"].numpy(), g["logits"]) < 1e-12
    assert rel_err(vals["mu"].numpy(), g["mu"]) < 1e-12
    for k in ("total_loss", "recon_loss", "kl_loss", "mutual_info", "mi_penalty", "collapse_penalty"):
        assert abs(float(vals[k]) - float(g["loss_" + k])) < 1e-12
    flat = O.tree_flatten({"encoder": ge, "decoder": gd})
    for n, t in flat.items():
        if name in FULL:
            assert rel_err(t.numpy(), g["grad/" + n]) < 1e-10 or float(np.abs(g["grad/" + n]).max()) == 0.0
        else:
            assert abs(float(t.norm()) - float(g["gradnorm/" + n])) <= 1e-10 * max(1.0, float(g["gradnorm/" + n]))


@pytest.mark.parametrize("name", FULL + ["default_b8"])
def test_fp32_oracle_within_1e4_of_fp64(name):
    g = load_golden(name)
    cfg, p, (vals, (ge, gd)) = _run(g, torch.float32)
    assert rel_err(vals["logits"].numpy(), g["logits"]) < 1e-4
    assert abs(float(vals["total_loss"]) - float(g["loss_total_loss"])) < 1e-4
    flat = O.tree_flatten({"encoder": ge, "decoder": gd})
    for n, t in flat.items():
        if name in FULL:
            ref = g["grad/" + n]
            assert float(np.abs(t.numpy() - ref).max()) <= 2e-4 * float(np.abs(ref).max()) + 1e-9, n


def test_lstm_layer_matches_torch_nn_lstm():
    """MLX nn.LSTM restatement vs torch.nn.LSTM: same gate order (i,f,g,o), single bias -> bias_hh = 0."""
    torch.manual_seed(0)
    B, T, D, H = 3, 9, 5, 7
    gen = torch.Generator().manual_seed(1)
    p = O._lstm_init(gen, D, H, torch.float64)
    x = torch.randn(B, T, D, dtype=torch.float64)
    ref = torch.nn.LSTM(D, H, batch_first=True).double()
    with torch.no_grad():
        ref.weight_ih_l0.copy_(p["Wx"]); ref.weight_hh_l0.copy_(p["Wh"])
        ref.bias_ih_l0.copy_(p["bias"]); ref.bias_hh_l0.zero_()
    want, (hn, cn) = ref(x)
    got_h, got_c = O.lstm_layer(p, x)
    assert torch.allclose(got_h, want, atol=1e-12)
    assert torch.allclose(got_c[:, -1], cn[0], atol=1e-12)


def test_finite_difference_gradients():
    g = load_golden("tiny")
    cfg = golden_cfg(g)
    p = golden_params(g)
    x = torch.as_tensor(g["x"]); c = torch.as_tensor(g["cond"]).double(); e = torch.as_tensor(g["eps"]).double()
    hyper = golden_hyper(g)
    # all-true coins: with greedy feedback the loss is only piecewise smooth in the decoder weights
    tf = np.ones(int(g["T"]), dtype=bool)
    vals, (ge, gd) = O.loss_and_grads(p, x, c, cfg.num_layers, e, tf, **hyper)
    grads = {"encoder": ge, "decoder": gd}
    rng = np.random.default_rng(0)
    for mod, name in [("encoder", "lstm_layer_0.Wh"), ("encoder", "fc_mu.weight"), ("encoder", "embedding.weight"),
                      ("decoder", "lstm_layer_1.Wx"), ("decoder", "fc_out.weight"), ("decoder", "lstm_layer_0.Wx")]:
        m, leaf = name.rsplit(".", 1)
        w = p[mod][m][leaf]
        for _ in range(3):
            idx = tuple(int(rng.integers(0, s)) for s in w.shape)
            h = 1e-6
            old = float(w[idx])
            w[idx] = old + h
            lp = float(O.complete_vae_loss(p, x, c, cfg.num_layers, e, tf, **hyper)["total_loss"])
            w[idx] = old - h
            lm = float(O.complete_vae_loss(p, x, c, cfg.num_layers, e, tf, **hyper)["total_loss"])
            w[idx] = old
            fd = (lp - lm) / (2 * h)
            an = float(grads[mod][m][leaf][idx])
            assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)) + 1e-8, (mod, name, idx, fd, an)


def test_reference_invariants():
    """component signs, tanh bounds, free-bits floor, dead decoder parameters (F1),
    collapse/mi penalty proportionality (F5)."""
    g = load_golden("tiny_sharp")
    cfg = golden_cfg(g)
    hyper = golden_hyper(g)
    p = golden_params(g)
    x = torch.as_tensor(g["x"]); c = torch.as_tensor(g["cond"]).double(); e = torch.as_tensor(g["eps"]).double()
    vals, (ge, gd) = O.loss_and_grads(p, x, c, cfg.num_layers, e, g["tf_mask"], **hyper)
    for k in ("total_loss", "recon_loss", "kl_loss", "collapse_penalty", "mutual_info", "mi_penalty"):
        assert float(vals[k]) >= 0
    assert float(vals["mu"].abs().max()) <= 2.0 and float(vals["logvar"].max()) <= 0.0 and float(vals["logvar"].min()) >= -2.0
    assert float(vals["kl_loss"]) >= hyper["free_bits"] - 1e-12
    assert abs(float(vals["collapse_penalty"]) / hyper["lambda_collapse"] - float(vals["mi_penalty"]) / hyper["lambda_mi"]) < 1e-12
    for mod in ("z_to_hidden", "condition_to_hidden"):
        for leaf in ("weight", "bias"):
            assert float(gd[mod][leaf].abs().max()) == 0.0
    H = cfg.hidden_dim
    for i in range(cfg.num_layers):
        assert float(gd[f"lstm_layer_{i}"]["Wh"].abs().max()) == 0.0
        assert float(gd[f"lstm_layer_{i}"]["Wx"][H:2 * H].abs().max()) == 0.0     # forget-gate rows
    # logits do not depend on z
    z2 = vals["z"] + 1.0
    l1 = O.decoder_forward(p["decoder"], vals["z"], c, cfg.num_layers, target_seq=x, tf_mask=g["tf_mask"])
    l2 = O.decoder_forward(p["decoder"], z2, c, cfg.num_layers, target_seq=x, tf_mask=g["tf_mask"])
    assert torch.equal(l1, l2)


def test_clip_is_a_noop_like_the_reference():
    g = load_golden("tiny")
    cfg, p, (vals, grads) = _run(g, torch.float64)
    big = tuple(O.tree_map(lambda t: t * 1e6, gr) for gr in grads)
    out = O.clip_gradients_reference(big, 1.0)
    assert out is big


def test_adam_is_mlx_adam_without_bias_correction():
    p = {"a": {"w": torch.tensor([1.0, -2.0], dtype=torch.float64)}}
    g = {"a": {"w": torch.tensor([0.5, 0.25], dtype=torch.float64)}}
    st = O.adam_init(p)
    newp, st = O.adam_update(p, g, st, lr=0.1)
    m = 0.1 * g["a"]["w"]; v = 0.001 * g["a"]["w"] ** 2
    want = p["a"]["w"] - 0.1 * m / (v.sqrt() + 1e-8)
    assert torch.allclose(newp["a"]["w"], want, atol=1e-15)


def test_golden_adam_step_consistent():
    g = load_golden("tiny")
    cfg = golden_cfg(g); p = golden_params(g)
    x = torch.as_tensor(g["x"]); c = torch.as_tensor(g["cond"]).double(); e = torch.as_tensor(g["eps"]).double()
    vals, grads, newp, _ = O.train_step(p, O.adam_init(p), x, c, cfg.num_layers, e, g["tf_mask"], float(g["lr"]), **golden_hyper(g))
    for n, t in O.tree_flatten(newp).items():
        assert rel_err(t.numpy(), g["newparam/" + n]) < 1e-12


def test_sampler_early_stop_and_shapes():
    cfg = O.Config(11, 8, 16, 8, 1, 2)
    p = O.init_params(cfg, seed=5)
    p["decoder"]["fc_out"]["bias"][cfg.end_token] = 50.0      # every row emits end at step 0
    c = torch.zeros(4, 1, dtype=torch.float64)
    toks = O.generate_with_temperature(p["decoder"], None if False else torch.zeros(4, 8).double(), c, 2, max_length=9)
    assert toks.shape == (4, 1) and bool((toks == cfg.end_token).all())
    toks = O.generate_with_temperature(p["decoder"], torch.zeros(4, 8).double(), c, 2, max_length=9, early_stopping=False)
    assert toks.shape == (4, 9)


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10 (the counter-based RNG the kernels restate)."""
    assert philox_ref.philox4x32(0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    full = 0xffffffffffffffff
    assert philox_ref.philox4x32(full, full, full) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)


def test_carry_state_decoder_is_a_stacked_lstm_with_initial_state():
    """carry_state=True (NOT the reference's behaviour): with teacher forcing everywhere the decoder is a
    standard stacked LSTM over [Emb[tok_{t-1}]; cond] with h0 = (z_to_hidden(z) + condition_to_hidden(cond)) / 2 in every
    layer and c0 = 0 (decoder.py:95-111) — cross-checked against torch.nn.LSTM; z now reaches the logits and
    z_to_hidden / Wh receive gradients (they are exactly zero in the reference mode, F1)."""
    cfg = O.Config(13, 8, 12, 6, 2, 3)
    p = O.init_params(cfg, seed=21, dtype=torch.float64)
    B, T = 5, 7
    x, cond, eps, _ = O.synthetic_batch(B, T, cfg, seed=22)
    xt, ct, z = torch.as_tensor(x).long(), torch.as_tensor(cond).double(), torch.as_tensor(eps).double()
    mask = np.ones(T, dtype=bool)
    lg = O.decoder_forward(p["decoder"], z, ct, cfg.num_layers, target_seq=xt, tf_mask=mask, carry_state=True)
    d = p["decoder"]
    ref = torch.nn.LSTM(cfg.embedding_dim + cfg.num_conditions, cfg.hidden_dim, num_layers=cfg.num_layers, batch_first=True).double()
    with torch.no_grad():
        for l in range(cfg.num_layers):
            getattr(ref, f"weight_ih_l{l}").copy_(d[f"lstm_layer_{l}"]["Wx"]); getattr(ref, f"weight_hh_l{l}").copy_(d[f"lstm_layer_{l}"]["Wh"])
            getattr(ref, f"bias_ih_l{l}").copy_(d[f"lstm_layer_{l}"]["bias"]); getattr(ref, f"bias_hh_l{l}").zero_()
    tok_in = torch.cat([torch.zeros(B, 1, dtype=torch.long), xt[:, :-1]], dim=1)
    inp = torch.cat([d["embedding"]["weight"][tok_in], ct.unsqueeze(1).expand(B, T, -1)], dim=2)
    h0, c0 = O.initialize_hidden_state(d, z, ct, cfg.num_layers)
    out, _ = ref(inp, (h0.contiguous(), c0.contiguous()))
    assert torch.allclose(lg, O.linear(d["fc_out"], out), atol=1e-12)
    # different z -> different logits; reference mode ignores z
    lg2 = O.decoder_forward(d, z + 1.0, ct, cfg.num_layers, target_seq=xt, tf_mask=mask, carry_state=True)
    assert float((lg2 - lg).abs().max()) > 1e-3
    a = O.decoder_forward(d, z, ct, cfg.num_layers, target_seq=xt, tf_mask=mask)
    b = O.decoder_forward(d, z + 1.0, ct, cfg.num_layers, target_seq=xt, tf_mask=mask)
    assert torch.equal(a, b)
    hyper = dict(beta=0.05, lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01, target_mi=4.85)
    _, (ge, gd) = O.loss_and_grads(p, xt, ct, cfg.num_layers, z, mask, carry_state=True, **hyper)
    assert float(gd["z_to_hidden"]["weight"].abs().max()) > 0 and float(gd["lstm_layer_1"]["Wh"].abs().max()) > 0
    assert float(gd["condition_to_hidden"]["bias"].abs().max()) > 0
    _, (ge0, _) = O.loss_and_grads(p, xt, ct, cfg.num_layers, z, mask, **hyper)
    # the reconstruction loss now reaches the encoder through z
    assert float((ge["fc_mu"]["weight"] - ge0["fc_mu"]["weight"]).abs().max()) > 1e-6
    # greedy sampling with state: first token equals the argmax of the first teacher-forced position
    toks = O.generate_with_temperature(d, z, ct, cfg.num_layers, max_length=4, early_stopping=False, carry_state=True)
    assert torch.equal(toks[:, 0], lg[:, 0].argmax(dim=1))
