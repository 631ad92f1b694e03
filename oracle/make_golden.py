#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the fp64 oracle (oracle/arcvae_oracle.py).

Test infrastructure.  The reference holds no golden vectors for this path and MLX
cannot run here (parity unpinned, see the oracle header), so these fixtures pin the
ORACLE: the CUDA path and the fp32 oracle are both compared against them.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Cases
  tiny        V11 E8 H16 L8 C1 NL2, B5 T7   — everything stored in full (params, grads, Adam)
  tiny_c2l3   V13 E8 H16 L8 C2 NL3, B4 T6   — num_conditions>1, 3 layers, in full
  default_b8  train.py default dims, B8 T16  — params regenerated from the seed; outputs in
              full, gradients as per-tensor norms + 64 seeded samples (keeps the file small)
  tiny_sharp / default_b8_sharp  decoder weights x4 / x6 after init so that greedy feedback
              (non-teacher-forced positions, the sampler) visits many different tokens
  loss_signs  the shapes of test_loss_signs.py:19-23 (B32 T120 V95 L128) — loss values only
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import arcvae_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
HYPER = dict(beta=0.05, lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01, target_mi=4.85)
LR = 2e-4


def sample_indices(numel, n=64, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, numel, size=min(n, numel))


def run_case(cfg, B, T, seed, tf_ratio, full, dec_scale=1.0):
    dt = torch.float64
    params = O.init_params(cfg, seed=seed, dtype=dt)
    if dec_scale != 1.0:   # sharpen the (otherwise nearly flat) random-init decoder so greedy chains vary
        params["decoder"] = O.tree_map(lambda t: t * dec_scale, params["decoder"])
    x, cond, eps, tf_mask = O.synthetic_batch(B, T, cfg, seed=seed + 1, tf_ratio=tf_ratio)
    xt, ct, et = torch.as_tensor(x), torch.as_tensor(cond).to(dt), torch.as_tensor(eps).to(dt)
    state = O.adam_init(params)
    vals, grads, new_params, new_state = O.train_step(params, state, xt, ct, cfg.num_layers, et, tf_mask, LR, **HYPER)
    _, dec_inputs = O.decoder_forward(params["decoder"], vals["z"], ct, cfg.num_layers, target_seq=xt,
                                      tf_mask=tf_mask, return_inputs=True)
    # evaluation forward: teacher_forcing_ratio = 0 (trainer.py:148, :459) -> all coins false
    ev = O.complete_vae_loss(params, xt, ct, cfg.num_layers, et, np.zeros(T, dtype=bool), return_logits=True, **HYPER)
    toks, margin = O.generate_with_temperature(params["decoder"], vals["z"], ct, cfg.num_layers, max_length=T,
                                               temperature=0.7, early_stopping=True, end_token=cfg.end_token,
                                               return_margin=True)
    d = {"cfg_" + k: np.asarray(v) for k, v in cfg.as_dict().items()}
    d.update(seed=np.asarray(seed), dec_scale=np.asarray(dec_scale), B=np.asarray(B), T=np.asarray(T), lr=np.asarray(LR),
             x=x, cond=cond, eps=eps, tf_mask=tf_mask, dec_inputs=dec_inputs.numpy().astype(np.int32),
             logits=vals["logits"].numpy(), mu=vals["mu"].numpy(), logvar=vals["logvar"].numpy(), z=vals["z"].numpy(),
             eval_logits=ev["logits"].detach().numpy(), eval_recon=np.asarray(float(ev["recon_loss"])),
             sample_tokens=toks.numpy().astype(np.int32), sample_margin=margin.detach().numpy())
    for k, v in HYPER.items():
        d["hyper_" + k] = np.asarray(v)
    for k in ("total_loss", "recon_loss", "kl_loss", "weighted_kl", "collapse_penalty", "prop_loss",
              "weighted_prop_loss", "mutual_info", "mi_penalty"):
        d["loss_" + k] = np.asarray(float(vals[k]))
    gflat = O.tree_flatten({"encoder": grads[0], "decoder": grads[1]})
    pflat = O.tree_flatten(params)
    nflat = O.tree_flatten(new_params)
    for n, g in gflat.items():
        g = g.numpy()
        if full:
            d["param/" + n] = pflat[n].numpy()
            d["grad/" + n] = g
            d["newparam/" + n] = nflat[n].numpy()
        else:
            idx = sample_indices(g.size, seed=sum(map(ord, n)))
            d["gradnorm/" + n] = np.asarray(np.linalg.norm(g))
            d["gradidx/" + n] = idx
            d["gradval/" + n] = g.reshape(-1)[idx]
            d["newparamval/" + n] = nflat[n].numpy().reshape(-1)[idx]
            d["paramsum/" + n] = np.asarray(pflat[n].numpy().sum())
    return d


def loss_signs_case():
    """Shapes of test_loss_signs.py:19-23 with seeded inputs (the script itself is unseeded)."""
    g = torch.Generator().manual_seed(5)
    B, T, V, L = 32, 120, 95, 128
    logits = torch.randn((B, T, V), generator=g, dtype=torch.float64)
    targets = torch.randint(0, V, (B, T), generator=g)
    mu = torch.randn((B, L), generator=g, dtype=torch.float64) * 0.1
    logvar = torch.randn((B, L), generator=g, dtype=torch.float64) * 0.1 - 1.0
    return dict(logits=logits.numpy().astype(np.float32), targets=targets.numpy().astype(np.int32),
                mu=mu.numpy(), logvar=logvar.numpy(),
                recon=np.asarray(float(O.reconstruction_loss(logits.float().double(), targets))),
                kl_fb0=np.asarray(float(O.kl_divergence(mu, logvar, free_bits=0.0))),
                kl_fb1=np.asarray(float(O.kl_divergence(mu, logvar, free_bits=1.0))),
                mi=np.asarray(float(O.mutual_information(mu, logvar))),
                collapse=np.asarray(float(O.posterior_collapse(mu, logvar, target_mi=4.85, weight=0.1))))


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "tiny.npz"),
                        **run_case(O.Config(11, 8, 16, 8, 1, 2), B=5, T=7, seed=11, tf_ratio=0.6, full=True))
    np.savez_compressed(os.path.join(OUT, "tiny_c2l3.npz"),
                        **run_case(O.Config(13, 8, 16, 8, 2, 3), B=4, T=6, seed=23, tf_ratio=0.5, full=True))
    np.savez_compressed(os.path.join(OUT, "default_b8.npz"),
                        **run_case(O.Config(), B=8, T=16, seed=67, tf_ratio=0.8, full=False))
    np.savez_compressed(os.path.join(OUT, "tiny_sharp.npz"),
                        **run_case(O.Config(11, 8, 16, 8, 1, 2), B=6, T=9, seed=31, tf_ratio=0.5, full=True, dec_scale=4.0))
    np.savez_compressed(os.path.join(OUT, "default_b8_sharp.npz"),
                        **run_case(O.Config(), B=8, T=16, seed=41, tf_ratio=0.5, full=False, dec_scale=6.0))
    np.savez_compressed(os.path.join(OUT, "loss_signs.npz"), **loss_signs_case())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
