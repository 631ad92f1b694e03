"""Shared helpers for the tests: golden fixtures, oracle/module construction."""
import os

import numpy as np
import torch

import arcvae_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HYPER_KEYS = ("beta", "lambda_prop", "lambda_collapse", "free_bits", "lambda_mi", "target_mi")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_cfg(g):
    return O.Config(int(g["cfg_vocab_size"]), int(g["cfg_embedding_dim"]), int(g["cfg_hidden_dim"]),
                    int(g["cfg_latent_dim"]), int(g["cfg_num_conditions"]), int(g["cfg_num_layers"]),
                    int(g["cfg_pad_token"]), int(g["cfg_end_token"]))


def golden_hyper(g):
    return {k: float(g["hyper_" + k]) for k in HYPER_KEYS}


def golden_params(g, dtype=torch.float64):
    """Full fixtures store the parameters; the default-dim ones regenerate them from the seed."""
    cfg = golden_cfg(g)
    if any(k.startswith("param/") for k in g):
        flat = {k[len("param/"):]: torch.as_tensor(v).to(dtype) for k, v in g.items() if k.startswith("param/")}
        return O.tree_unflatten(flat)
    p = O.init_params(cfg, seed=int(g["seed"]), dtype=torch.float64)
    s = float(g["dec_scale"])
    if s != 1.0:
        p["decoder"] = O.tree_map(lambda t: t * s, p["decoder"])
    for n, t in O.tree_flatten(p).items():      # guard: same torch RNG stream as when the fixture was made
        assert abs(float(t.sum()) - float(g["paramsum/" + n])) < 1e-9, n
    return O.tree_map(lambda t: t.to(dtype), p)


def model_kwargs(cfg):
    return dict(vocab_size=cfg.vocab_size, embedding_dim=cfg.embedding_dim, hidden_dim=cfg.hidden_dim,
                latent_dim=cfg.latent_dim, num_conditions=cfg.num_conditions, num_layers=cfg.num_layers)


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = max(float(np.abs(ref).max()), 1e-30)
    return float(np.abs(got - ref).max()) / den
