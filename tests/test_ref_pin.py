"""Pins the CPU oracle to the reference's OWN code (VERDICT r01 "Missing #1").

tests/golden/ref_*.npz hold what the unmodified reference sources (/root/reference: models/*.py, losses/*.py,
complete_vae_loss.py, trainer.py, mlx_data/dataloader.py) compute when executed under the ``mlx`` stand-in of
oracle/mlx_stub, in fp64 (generator: oracle/make_ref_golden.py -> oracle/ref_runner.py).  Here:
  * the oracle must reproduce every entry to 1e-10 relative (tokens / indices / batches exactly);
  * where the reference tree is mounted (the authoring container), the reference is re-run live and must equal the
    committed fixtures (stale-fixture guard);
  * the stand-in's MLX primitives are checked against independent implementations (torch.nn.LSTM, closed forms).
What stays unpinned: the MLX primitives themselves (no MLX binary exists in this image); see oracle/mlx_stub/mlx/__init__.py."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_runner as R  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _fixture(name):
    return dict(np.load(os.path.join(GOLDEN, f"ref_{name}.npz")))


@pytest.mark.parametrize("name", list(R.CASES))
def test_oracle_reproduces_the_reference_run(name):
    ref = _fixture(name)
    got = R.run_oracle(name)
    bad = R.compare(got, ref, rtol=1e-10)
    assert not bad, bad[:10]
    # the fixtures really cover the path: 12-key loss dict, both gradient trees, post-epoch weights and Adam moments
    for k in ("loss/total_loss", "loss/recon_loss", "loss/kl_loss", "loss/weighted_kl", "loss/collapse_penalty",
              "loss/prop_loss", "loss/weighted_prop_loss", "loss/mutual_info", "loss/mi_penalty", "loss/mu",
              "loss/logvar", "loss/z", "genc/lstm_layer_0.Wh", "gdec/fc_out.weight", "penc/fc_mu.weight",
              "pdec_opt/fc_out.weight.m", "tokens_early", "tokens_stop", "ds/batch0_mol", "epoch/loss"):
        assert k in ref, k
    assert bool(ref["clip_is_noop"]), "trainer.py:489-522 must leave the gradients untouched (F4)"
    assert ref["tokens_stop"].shape[1] == 1, "early stopping: every row emitted end_token at step 0"
    # F1: the decoder never sees z / Wh / z_to_hidden / condition_to_hidden -> exact-zero gradients in the REFERENCE run
    for k in ("gdec/z_to_hidden.weight", "gdec/condition_to_hidden.weight", "gdec/lstm_layer_0.Wh", "gdec/lstm_layer_1.Wh"):
        assert float(np.abs(ref[k]).max()) == 0.0, k


@pytest.mark.skipif(not R.reference_available(), reason="/root/reference is only mounted in the authoring container")
@pytest.mark.parametrize("name", list(R.CASES))
def test_committed_fixtures_equal_a_live_reference_run(name):
    live = R.run_reference(name)
    bad = R.compare(live, _fixture(name), rtol=1e-13)
    assert not bad, bad[:10]


# ---- the stand-in's primitives against independent implementations ---------------------------------------------------
@pytest.fixture()
def mlx():
    sys.path.insert(0, R.STUB)
    try:
        import mlx.core as mx
        import mlx.nn as nn
        import mlx.optimizers as optim
        mx.set_default_float(torch.float64)
        yield mx, nn, optim
    finally:
        sys.path.remove(R.STUB)


def test_stub_lstm_equals_torch_lstm(mlx):
    mx, nn, _ = mlx
    lstm = nn.LSTM(5, 7)
    ref = torch.nn.LSTM(5, 7, batch_first=True).double()
    with torch.no_grad():
        ref.weight_ih_l0.copy_(lstm.Wx); ref.weight_hh_l0.copy_(lstm.Wh)      # torch gate order is i,f,g,o as well
        ref.bias_ih_l0.copy_(lstm.bias); ref.bias_hh_l0.zero_()
    x = torch.randn(3, 9, 5, dtype=torch.float64)
    h, c = lstm(mx.array(x))
    out, (hn, cn) = ref(x)
    assert torch.allclose(h, out, atol=1e-12) and torch.allclose(c[:, -1], cn[0], atol=1e-12)
    # a fresh call on a length-1 sequence ignores Wh entirely (the decoder's situation, F1)
    h1, _ = lstm(mx.array(x[:, :1]))
    lstm.Wh = lstm.Wh * 0 + 123.0
    h2, _ = lstm(mx.array(x[:, :1]))
    assert torch.equal(h1, h2)


def test_stub_maximum_vjp_and_argmax_conventions(mlx):
    mx, _, _ = mlx
    a = mx.array([1.0, 2.0, 3.0]).requires_grad_(True)
    b = mx.array([2.0, 2.0, 2.0]).requires_grad_(True)
    mx.maximum(a, b).sum().backward()
    assert a.grad.tolist() == [0.0, 0.0, 1.0] and b.grad.tolist() == [1.0, 1.0, 0.0]      # tie -> second argument
    t = mx.array(0.5).requires_grad_(True)
    mx.maximum(0.0, 0.5 - t).backward()                # maximum(constant, x) at the kink x == 0 passes the gradient to x
    assert float(t.grad) == -1.0
    assert int(mx.argmax(mx.array([[1.0, 3.0, 3.0, 0.0]]), axis=1)) == 1                  # lowest index on ties


def test_stub_adam_has_no_bias_correction_and_modules_are_dicts(mlx):
    mx, nn, optim = mlx
    lin = nn.Linear(3, 2)
    assert isinstance(lin, dict) and set(lin.parameters()) == {"weight", "bias"} and lin.weight.shape == (2, 3)
    w0 = lin.weight.clone()
    g = {"weight": mx.ones_like(lin.weight) * 0.5, "bias": mx.zeros_like(lin.bias)}
    opt = optim.Adam(learning_rate=0.1)
    opt.update(lin, g)
    m, v = 0.1 * 0.5, 0.001 * 0.25
    assert torch.allclose(lin.weight, w0 - 0.1 * m / (np.sqrt(v) + 1e-8), atol=1e-12)
    assert abs(float(opt.state["weight"]["m"][0, 0]) - m) < 1e-15 and abs(float(opt.state["weight"]["v"][0, 0]) - v) < 1e-15

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(3, 2)
            self.depth = 7                             # plain attribute: not a parameter
    net = Net()
    assert "fc" in net and "depth" not in net and net.depth == 7
    loss, grads = mx.value_and_grad(lambda n, x: (n.fc(x) ** 2).sum(), argnums=[0])(net, mx.array([[1.0, 2.0, 3.0]]))
    assert type(grads[0]) is dict and type(grads[0]["fc"]) is dict and grads[0]["fc"]["weight"].shape == (2, 3)
    # every array sits one level below the top of the gradient dict: what makes trainer.py:502-507 sum nothing (F4)
    assert not any(isinstance(v, mx.array) for v in grads[0].values())


def test_reference_checkpoint_converts_to_the_flat_format(tmp_path, mlx):
    """tests/golden/ref_checkpoint_tiny.npz was written by the REFERENCE's save_checkpoint (trainer.py:577-603: pickled
    nested dicts of mx.array) at the end of the fixture epoch; tools/convert_mlx_checkpoint.py must turn it into the flat
    arcvae-flat-v1 file whose entries equal the post-epoch weights / Adam moments recorded in ref_tiny.npz."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("conv", os.path.join(ROOT, "tools", "convert_mlx_checkpoint.py"))
    conv = importlib.util.module_from_spec(spec); spec.loader.exec_module(conv)
    dst = str(tmp_path / "flat.npz")
    conv.main(os.path.join(GOLDEN, "ref_checkpoint_tiny.npz"), dst)
    flat = np.load(dst, allow_pickle=False)
    ref = _fixture("tiny")
    assert int(flat["epoch"]) == 3 and str(flat["format"]) == "arcvae-flat-v1"
    n = 0
    for tag, short in (("encoder", "penc"), ("decoder", "pdec")):
        for k in [k for k in ref if k.startswith(short + "/")]:
            name = k[len(short) + 1:]
            assert np.allclose(flat[f"{tag}/{name}"], ref[k], rtol=1e-6, atol=1e-7), (tag, name)
            n += 1
        for k in [k for k in ref if k.startswith(short + "_opt/")]:
            name, mom = k[len(short) + 5:].rsplit(".", 1)
            assert np.allclose(flat[f"{tag}_opt/{mom}/{name}"], ref[k], rtol=1e-6, atol=1e-12), (tag, name, mom)
            n += 1
    assert n > 80
    import json
    assert json.loads(str(flat["history_json"]))["epoch"] == [3.0]
