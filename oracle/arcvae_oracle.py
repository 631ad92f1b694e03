"""CPU oracle for the AR-CVAE hot path of Raiden-Makoto/MLX-VAE.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The shipped path
(``mlx-vae_b200/``) never routes through this file and has no CPU fallback.

PARITY: PINNED TO THE REFERENCE'S OWN PYTHON, MLX PRIMITIVES RESTATED.  The reference is pure Python over Apple
MLX (``requirements.txt:193`` ``mlx>=0.11.0``); MLX is not importable in this image and cannot be installed (no
network), the bundled dataset is absent (``.MISSING_LARGE_BLOBS:1``) and the reference holds no tests or golden
vectors for this path.  What pins this file:
  * ``tests/golden/ref_*.npz`` — outputs of the UNMODIFIED reference sources (models/*.py, losses/*.py,
    complete_vae_loss.py, trainer.py, mlx_data/dataloader.py) executed under the ``mlx`` stand-in of
    ``oracle/mlx_stub`` in fp64 (``oracle/make_ref_golden.py``); ``tests/test_ref_pin.py`` requires this oracle to
    reproduce every entry (forward values, the 12-key loss dict, both gradient trees, the no-op clip, a 3-batch
    trainer epoch with Adam, greedy tokens, dataset batches) to 1e-10, and re-runs the reference live where it is
    mounted;
  * ``tests/test_oracle.py`` — ``torch.nn.LSTM``, finite differences, fp32-vs-fp64, the reference's implied invariants,
    Random123 known answers for Philox.
What is still a restatement rather than a measurement: the MLX primitives the reference calls (``nn.LSTM``,
``nn.Linear``, ``nn.Embedding``, ``optim.Adam`` without bias correction, the ``mx.maximum`` VJP tie rule,
``mx.argmax`` tie-break, ``mx.value_and_grad`` over ``nn.Module`` trees) — SURVEY.md App. B; no MLX binary exists here.

Every function cites the reference file:line it restates (paths relative to
``/root/reference``).  The arithmetic is written with torch CPU tensors so the
same code runs in fp64 (golden vectors) and fp32 (timed CPU baseline), and so
that gradients come from reverse-mode autodiff over the same recorded graph,
as ``mx.value_and_grad`` does in ``trainer.py:292``.

Parameter tree (names are the reference's state-dict / gradient keys):
  encoder: embedding.weight [V,E]; lstm_layer_{i}.{Wx [4H,D], Wh [4H,H], bias [4H]};
           condition_fc, fc_mu, fc_logvar_hidden, fc_logvar .{weight [out,in], bias}
  decoder: z_to_hidden, condition_to_hidden, embedding, lstm_layer_{i}, fc_out
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

Tree = Dict[str, Dict[str, torch.Tensor]]


# --------------------------------------------------------------------------- #
# configuration + MLX-style initialisation
# --------------------------------------------------------------------------- #
class Config:
    """Model dims.  Defaults are ``train.py:26-31`` (the argparse defaults)."""

    def __init__(self, vocab_size=80, embedding_dim=128, hidden_dim=256, latent_dim=128,
                 num_conditions=1, num_layers=2, pad_token=0, end_token=2):
        self.vocab_size = vocab_size
        self.embedding_dim = embedding_dim
        self.hidden_dim = hidden_dim
        self.latent_dim = latent_dim
        self.num_conditions = num_conditions
        self.num_layers = num_layers
        self.pad_token = pad_token      # decoder.py:25
        self.end_token = end_token      # decoder.py:26

    def as_dict(self):
        return dict(self.__dict__)


def _uniform(gen, shape, bound, dtype):
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2.0 - 1.0) * bound).to(dtype)


def _linear_init(gen, out_dim, in_dim, dtype):
    # MLX nn.Linear: weight [out,in], bias [out], both U(-1/sqrt(in), 1/sqrt(in)).
    k = 1.0 / math.sqrt(in_dim)
    return {"weight": _uniform(gen, (out_dim, in_dim), k, dtype), "bias": _uniform(gen, (out_dim,), k, dtype)}


def _lstm_init(gen, in_dim, hid, dtype):
    # MLX nn.LSTM: Wx [4H,D], Wh [4H,H], bias [4H], all U(-1/sqrt(H), 1/sqrt(H)).
    k = 1.0 / math.sqrt(hid)
    return {"Wx": _uniform(gen, (4 * hid, in_dim), k, dtype),
            "Wh": _uniform(gen, (4 * hid, hid), k, dtype),
            "bias": _uniform(gen, (4 * hid,), k, dtype)}


def _embedding_init(gen, vocab, dim, dtype):
    # MLX nn.Embedding: weight [V,E] ~ N(0, 1/E).
    w = torch.randn((vocab, dim), generator=gen, dtype=torch.float64) * math.sqrt(1.0 / dim)
    return {"weight": w.to(dtype)}


def init_encoder_params(cfg: Config, gen: torch.Generator, dtype=torch.float64) -> Tree:
    """models/encoder.py:46-74 (module construction order kept)."""
    p: Tree = {"embedding": _embedding_init(gen, cfg.vocab_size, cfg.embedding_dim, dtype)}
    for i in range(cfg.num_layers):
        d = cfg.embedding_dim if i == 0 else cfg.hidden_dim
        p[f"lstm_layer_{i}"] = _lstm_init(gen, d, cfg.hidden_dim, dtype)
    p["condition_fc"] = _linear_init(gen, cfg.hidden_dim, cfg.num_conditions, dtype)
    comb = 2 * cfg.hidden_dim
    p["fc_mu"] = _linear_init(gen, cfg.latent_dim, comb, dtype)
    p["fc_logvar_hidden"] = _linear_init(gen, comb, comb, dtype)
    p["fc_logvar"] = _linear_init(gen, cfg.latent_dim, comb, dtype)
    p["fc_logvar"]["bias"] = torch.full((cfg.latent_dim,), 0.35, dtype=dtype)   # encoder.py:71-74
    return p


def init_decoder_params(cfg: Config, gen: torch.Generator, dtype=torch.float64) -> Tree:
    """models/decoder.py:51-73."""
    p: Tree = {"z_to_hidden": _linear_init(gen, cfg.hidden_dim, cfg.latent_dim, dtype),
               "condition_to_hidden": _linear_init(gen, cfg.hidden_dim, cfg.num_conditions, dtype),
               "embedding": _embedding_init(gen, cfg.vocab_size, cfg.embedding_dim, dtype)}
    for i in range(cfg.num_layers):
        d = cfg.embedding_dim + cfg.num_conditions if i == 0 else cfg.hidden_dim
        p[f"lstm_layer_{i}"] = _lstm_init(gen, d, cfg.hidden_dim, dtype)
    p["fc_out"] = _linear_init(gen, cfg.vocab_size, cfg.hidden_dim, dtype)
    return p


def init_params(cfg: Config, seed: int = 67, dtype=torch.float64) -> Dict[str, Tree]:
    gen = torch.Generator().manual_seed(seed)
    return {"encoder": init_encoder_params(cfg, gen, dtype), "decoder": init_decoder_params(cfg, gen, dtype)}


def tree_map(fn, tree):
    return {k: (tree_map(fn, v) if isinstance(v, dict) else fn(v)) for k, v in tree.items()}


def tree_flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        name = f"{prefix}{k}"
        if isinstance(v, dict):
            out.update(tree_flatten(v, name + "."))
        else:
            out[name] = v
    return out


def tree_unflatten(flat):
    out: dict = {}
    for name, v in flat.items():
        parts = name.split(".")
        d = out
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return out


# --------------------------------------------------------------------------- #
# MLX primitives restated
# --------------------------------------------------------------------------- #
def linear(p, v):
    """MLX nn.Linear.__call__: addmm(bias, x, weight.T)."""
    return v @ p["weight"].T + p["bias"]


def lstm_layer(p, x, hidden=None, cell=None):
    """MLX ``nn.LSTM.__call__(x, hidden=None, cell=None)`` restated.

    ``x = addmm(bias, x, Wx.T)`` for all positions at once, then a loop over the
    sequence axis; ``hidden @ Wh.T`` is added only when ``hidden is not None``;
    gate order i, f, g, o; ``cell = f*cell + i*g`` when ``cell is not None`` else
    ``i*g``; returns ``(all_hidden, all_cell)`` each [B, T, H].
    Call sites: models/encoder.py:99-101, models/decoder.py:165-168.
    """
    H = p["Wh"].shape[1]
    proj = x @ p["Wx"].T + p["bias"]
    all_h, all_c = [], []
    for idx in range(proj.shape[-2]):
        ifgo = proj[..., idx, :]
        if hidden is not None:
            ifgo = ifgo + hidden @ p["Wh"].T
        i, f, g, o = torch.split(ifgo, H, dim=-1)
        i = torch.sigmoid(i)
        f = torch.sigmoid(f)
        g = torch.tanh(g)
        o = torch.sigmoid(o)
        cell = f * cell + i * g if cell is not None else i * g
        hidden = o * torch.tanh(cell)
        all_c.append(cell)
        all_h.append(hidden)
    return torch.stack(all_h, dim=-2), torch.stack(all_c, dim=-2)


# --------------------------------------------------------------------------- #
# model
# --------------------------------------------------------------------------- #
def encoder_forward(p: Tree, x: torch.Tensor, conditions: torch.Tensor, num_layers: int):
    """models/encoder.py:76-132.  x [B,T] integer, conditions [B,C] -> (mu, logvar) [B,L]."""
    embedded = p["embedding"]["weight"][x.long()]                      # :93
    output = embedded
    for i in range(num_layers):                                        # :98-101
        output, _hidden = lstm_layer(p[f"lstm_layer_{i}"], output)
    final_hidden = output[:, -1, :]                                    # :106 (through padding, F2)
    condition_repr = linear(p["condition_fc"], conditions)             # :109
    combined = torch.cat([final_hidden, condition_repr], dim=1)        # :112
    mu_raw = linear(p["fc_mu"], combined)                              # :115
    logvar_hidden = torch.tanh(linear(p["fc_logvar_hidden"], combined))  # :117
    logvar_raw = linear(p["fc_logvar"], logvar_hidden)                 # :118
    mu = torch.tanh(mu_raw / 2.0) * 2.0                                # :126
    logvar = torch.tanh(logvar_raw / 2.0) * 1.0 - 1.0                  # :130
    return mu, logvar


def reparameterize(mu, logvar, eps):
    """models/encoder.py:147-153 with the N(0,1) draw injected (``eps``)."""
    std = torch.exp(0.5 * logvar)
    return mu + eps * std


def initialize_hidden_state(p: Tree, z, conditions, num_layers: int):
    """models/decoder.py:76-111 (computed by the reference, never consumed: F1)."""
    hidden_init = (linear(p["z_to_hidden"], z) + linear(p["condition_to_hidden"], conditions)) / 2.0
    hidden = hidden_init.unsqueeze(0).repeat(num_layers, 1, 1)
    return hidden, torch.zeros_like(hidden)


def _decoder_step(p: Tree, current_token, conditions, num_layers: int, state=None):
    """models/decoder.py:154-175 — one position.  ``state is None``: LSTM layers called WITHOUT state, as the reference
    does (F1).  ``state = (h, c)`` (lists of [B,H] per layer): the ``carry_state=True`` extension (the
    evidently intended semantics of decoder.py:95-99, :143): every layer is called with its own (hidden, cell) and the
    state is updated in place."""
    embedded = p["embedding"]["weight"][current_token.long()]          # :154
    lstm_input = torch.cat([embedded, conditions], dim=1).unsqueeze(1)  # :157-160
    output = lstm_input
    for i in range(num_layers):                                        # :165-168
        if state is None:
            output, _hidden = lstm_layer(p[f"lstm_layer_{i}"], output)
        else:
            output, cells = lstm_layer(p[f"lstm_layer_{i}"], output, hidden=state[0][i], cell=state[1][i])
            state[0][i], state[1][i] = output[:, -1, :], cells[:, -1, :]
    return linear(p["fc_out"], output[:, 0, :])                        # :172-175


def decoder_forward(p: Tree, z, conditions, num_layers: int, target_seq=None, max_length: int = 80,
                    tf_mask: Optional[Sequence[bool]] = None, return_inputs: bool = False, carry_state: bool = False):
    """models/decoder.py:113-190.

    ``tf_mask[t]`` replaces the host coin ``np.random.rand() < teacher_forcing_ratio``
    of ``decoder.py:180`` (one coin per position for the whole batch, F7).  With
    ``target_seq is None`` every coin is false, as in the reference.
    ``carry_state=True`` (NOT the reference's behaviour): the state built by ``initialize_hidden_state`` is consumed and
    carried from position to position.
    """
    B = z.shape[0]
    T = target_seq.shape[1] if target_seq is not None else max_length   # :137-140
    hidden, cell = initialize_hidden_state(p, z, conditions, num_layers)   # :143 dead value unless carry_state
    state = ([hidden[i] for i in range(num_layers)], [cell[i] for i in range(num_layers)]) if carry_state else None
    current = torch.zeros((B,), dtype=torch.long)                      # :146 start token 0
    outs, inputs = [], []
    for t in range(T):
        inputs.append(current)
        logits = _decoder_step(p, current, conditions, num_layers, state)
        outs.append(logits)
        use_tf = target_seq is not None and tf_mask is not None and bool(tf_mask[t])
        if use_tf:
            current = target_seq[:, t].long()                          # :182
        else:
            current = torch.argmax(logits.detach(), dim=1)             # :185 (lowest index on ties)
    out = torch.stack(outs, dim=1)                                     # :188
    if return_inputs:
        return out, torch.stack(inputs, dim=1)
    return out


def generate_with_temperature(p: Tree, z, conditions, num_layers: int, max_length: int = 80,
                              temperature: float = 1.0, early_stopping: bool = True, end_token: int = 2,
                              return_margin: bool = False, carry_state: bool = False):
    """models/decoder_sampling.py:48-128.  Returns tokens [B, t_stop<=max_length] (int64).

    "Sampling" is argmax(softmax(logits/temperature)) (:110-117, F8).  With
    ``return_margin`` also returns the per-position top1-top2 logit gap so tests can
    classify near-tie disagreements instead of hiding them.
    """
    B = z.shape[0]
    hidden, cell = initialize_hidden_state(p, z, conditions, num_layers)   # :75 dead value unless carry_state
    state = ([hidden[i] for i in range(num_layers)], [cell[i] for i in range(num_layers)]) if carry_state else None
    current = torch.zeros((B,), dtype=torch.long)                      # :78
    ended = torch.zeros((B,), dtype=torch.bool)                        # :81
    toks, margins = [], []
    for _t in range(max_length):
        if early_stopping and bool(ended.all()):                       # :87-88
            break
        logits = _decoder_step(p, current, conditions, num_layers, state)
        probs = torch.softmax(logits / temperature, dim=1)             # :110-113
        current = torch.argmax(probs, dim=1)                           # :117
        toks.append(current)
        if return_margin:
            top2 = torch.topk(logits, 2, dim=1).values
            margins.append(top2[:, 0] - top2[:, 1])
        ended = ended | (current == end_token)                         # :122-123
    tokens = torch.stack(toks, dim=1) if toks else torch.zeros((B, 0), dtype=torch.long)
    if return_margin:
        return tokens, (torch.stack(margins, dim=1) if margins else torch.zeros((B, 0)))
    return tokens


def vae_forward(params, x, conditions, num_layers, eps, target_seq=None, tf_mask=None):
    """models/vae.py:63-99."""
    mu, logvar = encoder_forward(params["encoder"], x, conditions, num_layers)
    z = reparameterize(mu, logvar, eps)
    logits = decoder_forward(params["decoder"], z, conditions, num_layers, target_seq=target_seq, tf_mask=tf_mask)
    return logits, mu, logvar, z


# --------------------------------------------------------------------------- #
# losses
# --------------------------------------------------------------------------- #
def reconstruction_loss(logits, targets, reduction="mean"):
    """losses/recon.py:29-64 — max-shifted log-softmax, gather, UNMASKED mean (F3)."""
    V = logits.shape[-1]
    lf = logits.reshape(-1, V)
    tf = targets.reshape(-1).long()
    m = lf.max(dim=1, keepdim=True).values
    st = lf - m
    log_softmax = st - torch.log(torch.exp(st).sum(dim=1, keepdim=True))
    ce = -torch.take_along_dim(log_softmax, tf.reshape(-1, 1), dim=1).reshape(-1)
    if reduction == "mean":
        return ce.mean()
    if reduction == "sum":
        return ce.sum()
    return ce


def _mx_maximum(a, b):
    """mx.maximum with MLX's VJP convention: cotangent to ``a`` where a > b, else to ``b``."""
    if not torch.is_tensor(b):
        b = torch.as_tensor(b, dtype=a.dtype)
    if not torch.is_tensor(a):
        a = torch.as_tensor(a, dtype=b.dtype)
    return torch.where(a > b, a, b)


def kl_divergence(mu, logvar, reduction="mean", free_bits: float = 0.0):
    """losses/kl.py:35-66."""
    latent_dim = mu.shape[1]
    mu = torch.clamp(mu, -3.0, 3.0)                                     # :39
    logvar = torch.clamp(logvar, -6.0, 3.0)                            # :40
    var = torch.exp(logvar)
    kl_per_dim = -0.5 * (1.0 + logvar - mu * mu - var)                 # :48
    kl_per_dim = _mx_maximum(kl_per_dim, 0.0)                          # :51
    if free_bits > 0.0:                                                # :54-56
        kl_per_dim = _mx_maximum(kl_per_dim, free_bits / latent_dim)
    per_sample = kl_per_dim.sum(dim=1)                                 # :59
    if reduction == "mean":
        return per_sample.mean()
    if reduction == "sum":
        return per_sample.sum()
    return per_sample


def mutual_information(mu, logvar):
    """losses/info.py:23-50 (no epsilon inside log(mean_var), :38)."""
    mu = torch.clamp(mu, -3.0, 3.0)
    logvar = torch.clamp(logvar, -6.0, 3.0)
    var = torch.exp(logvar)
    kl_per_sample = -0.5 * (1.0 + logvar - mu * mu - var).sum(dim=1)   # :32
    mean_kl = kl_per_sample.mean()
    mean_mu = mu.mean(dim=0)
    mean_var = var.mean(dim=0)
    mean_logvar = torch.log(mean_var)
    agg_kl = -0.5 * (1.0 + mean_logvar - mean_mu * mean_mu - mean_var).sum()  # :41
    return _mx_maximum(mean_kl - agg_kl, 0.0)                          # :46-48


def posterior_collapse(mu, logvar, target_mi: float = 4.85, weight: float = 0.1):
    """losses/info.py:73-78."""
    mi = mutual_information(mu, logvar)
    return weight * _mx_maximum(torch.zeros((), dtype=mi.dtype), target_mi - mi)


def property_prediction_loss(z, predicted_properties, target_properties, property_scales=None, reduction="mean"):
    """losses/prop.py:5-40 (dead on the training path, F10)."""
    mse = (predicted_properties - target_properties) ** 2
    if property_scales is not None:
        mse = mse / (property_scales ** 2 + 1e-8)
    if reduction == "mean":
        return mse.mean()
    if reduction == "sum":
        return mse.sum()
    return mse


def complete_vae_loss(params, x, conditions, num_layers, eps, tf_mask, beta=0.4, lambda_prop=0.1,
                      lambda_collapse=0.01, free_bits=0.5, lambda_mi=0.0, target_mi=4.85, return_logits=False,
                      carry_state=False):
    """complete_vae_loss.py:37-99 with ``property_predictor=None`` (train.py:186)."""
    mu, logvar = encoder_forward(params["encoder"], x, conditions, num_layers)       # :38
    z = reparameterize(mu, logvar, eps)                                              # :39
    logits = decoder_forward(params["decoder"], z, conditions, num_layers, target_seq=x, tf_mask=tf_mask,
                             carry_state=carry_state)                                # :42
    recon = reconstruction_loss(logits, x)                                           # :45
    kl = kl_divergence(mu, logvar, free_bits=free_bits)                              # :48
    collapse = posterior_collapse(mu, logvar, weight=lambda_collapse)                # :51 (target 4.85 default)
    mi = mutual_information(mu, logvar)                                              # :59
    mi_penalty = lambda_mi * _mx_maximum(torch.zeros((), dtype=mi.dtype), target_mi - mi)  # :60
    prop = torch.zeros((), dtype=mu.dtype)                                           # :67
    total = recon + beta * kl + collapse + lambda_prop * prop + mi_penalty          # :76-82
    out = {"total_loss": total, "recon_loss": recon, "kl_loss": kl, "weighted_kl": beta * kl,
           "collapse_penalty": collapse, "prop_loss": prop, "weighted_prop_loss": lambda_prop * prop,
           "mutual_info": mi, "mi_penalty": mi_penalty, "mu": mu, "logvar": logvar, "z": z}
    if return_logits:
        out["logits"] = logits
    return out


# --------------------------------------------------------------------------- #
# trainer step
# --------------------------------------------------------------------------- #
def loss_and_grads(params, x, conditions, num_layers, eps, tf_mask, **hyper):
    """trainer.py:267-305 — ``mx.value_and_grad(model_loss_fn, argnums=[0,1])``.

    Returns (loss dict with detached values incl. logits, (enc_grads, dec_grads)) where the
    grads are nested dicts keyed like the parameter tree; unused parameters get zeros.
    """
    leaves = tree_map(lambda t: t.detach().clone().requires_grad_(True), params)
    out = complete_vae_loss(leaves, x, conditions, num_layers, eps, tf_mask, return_logits=True, **hyper)
    flat = tree_flatten(leaves)
    names = list(flat.keys())
    gs = torch.autograd.grad(out["total_loss"], [flat[n] for n in names], allow_unused=True)
    gflat = {n: (g if g is not None else torch.zeros_like(flat[n])) for n, g in zip(names, gs)}
    gtree = tree_unflatten(gflat)
    vals = {k: v.detach() for k, v in out.items()}
    return vals, (gtree["encoder"], gtree["decoder"])


def clip_gradients_reference(grads, max_norm: float = 1.0):
    """trainer.py:489-522.  Only arrays at the TOP level of each grad dict are summed; every
    leaf sits one level down, so the norm is 0 and nothing is scaled (F4)."""
    total = 0.0
    for g in grads:
        if isinstance(g, dict):
            for v in g.values():
                if torch.is_tensor(v):
                    total += float((v * v).sum())
    norm = math.sqrt(total)
    if norm > max_norm:
        scale = max_norm / (norm + 1e-8)
        return tuple({k: (v * scale if torch.is_tensor(v) else v) for k, v in g.items()} for g in grads)
    return grads


def adam_init(params):
    return {"m": tree_map(torch.zeros_like, params), "v": tree_map(torch.zeros_like, params)}


def adam_update(params, grads, state, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """MLX ``optim.Adam`` (trainer.py:75-76, :320-324): no bias correction.
    m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2 ; p = p - lr*m/(sqrt(v)+eps)."""
    fp, fg, fm, fv = (tree_flatten(t) for t in (params, grads, state["m"], state["v"]))
    np_, nm, nv = {}, {}, {}
    for n in fp:
        m = b1 * fm[n] + (1.0 - b1) * fg[n]
        v = b2 * fv[n] + (1.0 - b2) * fg[n] * fg[n]
        np_[n] = fp[n] - lr * m / (torch.sqrt(v) + eps)
        nm[n], nv[n] = m, v
    return tree_unflatten(np_), {"m": tree_unflatten(nm), "v": tree_unflatten(nv)}


def train_step(params, opt_state, x, conditions, num_layers, eps, tf_mask, lr, grad_clip=1.0, **hyper):
    """trainer.py:305-333: value_and_grad, (no-op) clip, two Adam updates."""
    vals, grads = loss_and_grads(params, x, conditions, num_layers, eps, tf_mask, **hyper)
    if grad_clip > 0:
        grads = clip_gradients_reference(grads, grad_clip)
    new_enc, st_enc = adam_update(params["encoder"], grads[0],
                                  {"m": opt_state["m"]["encoder"], "v": opt_state["v"]["encoder"]}, lr)
    new_dec, st_dec = adam_update(params["decoder"], grads[1],
                                  {"m": opt_state["m"]["decoder"], "v": opt_state["v"]["decoder"]}, lr)
    new_params = {"encoder": new_enc, "decoder": new_dec}
    new_state = {"m": {"encoder": st_enc["m"], "decoder": st_dec["m"]},
                 "v": {"encoder": st_enc["v"], "decoder": st_dec["v"]}}
    return vals, grads, new_params, new_state


# --------------------------------------------------------------------------- #
# synthetic SELFIES-shaped data
# --------------------------------------------------------------------------- #
def synthetic_batch(B: int, T: int, cfg: Config, seed: int = 67, tf_ratio: float = 0.9):
    """Lengths U{min(16,T)..T}; tokens U{3..V-1}; end token 2 at len-1; pad 0 after
    (mlx_data/dataloader.py:76-79); cond ~ N(0,1) (z-scored TPSA, dataloader.py:63-65)."""
    rng = np.random.default_rng(seed)
    lo = min(16, T)
    lens = rng.integers(lo, T + 1, size=B)
    x = np.zeros((B, T), dtype=np.int32)
    body = rng.integers(3, cfg.vocab_size, size=(B, T))
    for b in range(B):
        n = int(lens[b])
        x[b, : n - 1] = body[b, : n - 1]
        x[b, n - 1] = cfg.end_token
    cond = rng.standard_normal((B, cfg.num_conditions)).astype(np.float32)
    eps = rng.standard_normal((B, cfg.latent_dim)).astype(np.float32)
    tf_mask = rng.random(T) < tf_ratio
    return x, cond, eps, tf_mask


def to_torch(params_np, dtype):
    return tree_map(lambda a: torch.as_tensor(np.asarray(a)).to(dtype), params_np)
