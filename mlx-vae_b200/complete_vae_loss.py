"""complete_vae_loss — drop-in for complete_vae_loss.py of the reference, plus the gradient entry point that
replaces ``mx.value_and_grad(model_loss_fn, argnums=[0, 1])`` (trainer.py:292)."""
from __future__ import annotations

from typing import Optional

import os

import torch

from . import _lib
from .losses._fused import fused_loss, make_hyper


def _run(encoder, decoder, property_predictor, x, conditions, beta, lambda_prop, lambda_collapse,
         teacher_forcing_ratio, free_bits, lambda_mi, target_mi, eps, tf_mask, seed, pad_mask, backward, allreduce,
         return_logits=False, backward_hooks=None, eps_offset=0, ce_world=1, info=None):
    if property_predictor is not None:
        # the reference would raise TypeError here (complete_vae_loss.py:63-67 vs losses/prop.py:5-11, F10)
        raise NotImplementedError("property_predictor must be None, as in train.py:186")
    hp = make_hyper(beta, lambda_prop, lambda_collapse, free_bits, lambda_mi, target_mi, 4.85, pad_mask)
    targets = encoder._tokens(x)
    B_, T_ = targets.shape
    fuse_ce = (backward and not return_logits and not pad_mask and os.environ.get("ARCVAE_NO_FUSED_CE") is None
               and decoder.ce_supported(B_))
    if fuse_ce and os.environ.get("ARCVAE_ONE_STREAM") is None:
        return _run_two_streams(encoder, decoder, x, conditions, targets, hp, teacher_forcing_ratio, eps, tf_mask, seed,
                                eps_offset, info, allreduce, backward_hooks, ce_world)
    mu, logvar = encoder(x, conditions)                                                         # :38
    z_in = None
    if getattr(decoder, "carry_state", False):
        # extension mode: the decoder consumes z (reference mode: z is dead, F1); same Philox stream as the loss kernel
        z_in = encoder.reparameterize(mu, logvar, eps, seed=seed, offset=eps_offset)            # :39
    if fuse_ce:
        # training step on the fused bf16 path: fc_out + cross-entropy + d logits + greedy feedback in ONE GEMM epilogue;
        # the fp32 logits never reach HBM.  Unmasked mean over the GLOBAL batch (losses/recon.py:59-60): under data
        # parallelism the shards are equal (trainer contract), so the global position count is world x B x T.
        L_ = mu.shape[1]
        stats = torch.zeros(2 * L_ + 6, dtype=torch.float64, device=mu.device)
        decoder.forward_ce(conditions, targets, stats[2 * L_ + 2:2 * L_ + 3], 1.0 / (float(ce_world) * B_ * T_),
                           teacher_forcing_ratio, tf_mask=tf_mask)                              # :42, :45
        out = fused_loss(None, None, mu, logvar, hp, eps=eps, seed=seed, offset=eps_offset, pad_token=decoder.pad_token,
                         want_grads=True, want_z=True, allreduce=allreduce, stats=stats, ce_T=T_)   # :39, :48-82
        logits = None
    else:
        logits = decoder(z_in, conditions, target_seq=x, teacher_forcing_ratio=teacher_forcing_ratio,
                         tf_mask=tf_mask)                                                       # :42
        out = fused_loss(logits, targets, mu, logvar, hp, eps=eps, seed=seed, offset=eps_offset, pad_token=decoder.pad_token,
                         want_grads=backward, want_z=True, inplace_dlogits=backward and not return_logits,
                         allreduce=allreduce)                                                   # :39, :45-82
    d = {k: out.losses[i] for i, k in enumerate(_lib.LOSS_KEYS)}                                # :86-99
    d.update(mu=mu, logvar=logvar, z=out.z)
    # under data parallelism the unfused path reports recon = local CE / global tokens (a partial the trainer sums through
    # the gradient all-reduce); on the fused-CE path the CE sum is part of the all-reduced statistics: already global
    if info is not None:
        info["recon_is_global"] = bool(fuse_ce and allreduce is not None)
        info["fused_ce"] = bool(fuse_ce)
    if return_logits:
        d["logits"] = logits
    if allreduce is not None and not backward:
        # forward-only under data parallelism: the kernel reports local_CE / global_tokens (see trainer.train_step)
        recon = d["recon_loss"].clone()
        allreduce(recon.reshape(1))
        d["recon_loss"] = recon
        d["total_loss"] = recon + d["weighted_kl"] + d["collapse_penalty"] + d["weighted_prop_loss"] + d["mi_penalty"]
    if backward:
        dz = decoder.backward(out.dlogits)
        if dz is not None:                      # carry_state: the reconstruction loss reaches the encoder through z
            B_, L_ = mu.shape
            _lib.check(_lib.load().arcvae_reparam_backward(dz.data_ptr(), z_in.data_ptr(), mu.data_ptr(), B_, L_,
                                                           out.dmu.data_ptr(), out.dlogvar.data_ptr(), _lib.stream_ptr()))
        if backward_hooks is not None:
            backward_hooks.after_decoder_backward(decoder, d, bool(fuse_ce and allreduce is not None))
        encoder.backward(out.dmu, out.dlogvar)
    return d


def _run_two_streams(encoder, decoder, x, conditions, targets, hp, teacher_forcing_ratio, eps, tf_mask, seed, eps_offset,
                     info, allreduce, backward_hooks, ce_world):
    """Training step of the fused bf16 path with the ENCODER chain on a second (high-priority) stream.

    In the reference mode the decoder never sees z (F1) and, with the cross-entropy in the fc_out epilogue, its reverse
    pass needs nothing from the loss kernel: decoder forward + backward are independent of the whole encoder chain; only
    the loss kernel joins them (it reads the CE sum).  The persistent cluster recurrence holds 128 of the 148 SMs for
    ~4.5 ms at low occupancy; the decoder's kernels fill the 20 SMs it leaves (measured 9.41 -> 8.94 ms per step).
    Under data parallelism the statistics all-reduce is enqueued on the encoder stream and the decoder's gradient
    all-reduce starts as soon as its reverse pass is enqueued, as before.  ARCVAE_ONE_STREAM=1 restores the serial order."""
    B_, T_ = targets.shape
    L_ = encoder.latent_dim
    main = torch.cuda.current_stream()
    hi = getattr(encoder, "_hi_stream", None)
    if hi is None:
        hi = encoder._hi_stream = torch.cuda.Stream(device=encoder.device, priority=-1)
    # every tensor of the step is allocated HERE, on the caller's stream: the side stream only runs kernels on them and is
    # bracketed by hi.wait_stream(main) / main.wait_stream(hi), so the caching allocator never sees a cross-stream block
    # (record_stream() defers reuse and ends in sporadic cudaMalloc stalls of the whole device)
    dev = decoder.device
    stats = torch.zeros(2 * L_ + 6, dtype=torch.float64, device=dev)
    cond32 = encoder._f32(conditions)
    mu = torch.empty((B_, L_), dtype=torch.float32, device=dev)
    logvar = torch.empty_like(mu)
    bufs = {"losses": torch.empty(len(_lib.LOSS_KEYS), dtype=torch.float32, device=dev), "dmu": torch.empty_like(mu),
            "dlogvar": torch.empty_like(mu), "z": torch.empty_like(mu)}
    if eps is not None:
        eps = eps.float().contiguous()
    hi.wait_stream(main)
    with torch.cuda.stream(hi):
        encoder(targets, cond32, out=(mu, logvar))                                              # :38
    decoder.forward_ce(conditions, targets, stats[2 * L_ + 2:2 * L_ + 3], 1.0 / (float(ce_world) * B_ * T_),
                       teacher_forcing_ratio, tf_mask=tf_mask)                                  # :42, :45
    ce_done = main.record_event()
    with torch.cuda.stream(hi):
        hi.wait_event(ce_done)
        out = fused_loss(None, None, mu, logvar, hp, eps=eps, seed=seed, offset=eps_offset, pad_token=decoder.pad_token,
                         want_grads=True, want_z=True, allreduce=allreduce, stats=stats, ce_T=T_, bufs=bufs)   # :39, :48-82
        loss_done = hi.record_event()
        encoder.backward(out.dmu, out.dlogvar)
    d = {k: out.losses[i] for i, k in enumerate(_lib.LOSS_KEYS)}                                # :86-99
    d.update(mu=mu, logvar=logvar, z=out.z)
    if info is not None:
        info["recon_is_global"], info["fused_ce"] = allreduce is not None, True
    decoder.backward(None)
    if backward_hooks is not None:
        main.wait_event(loss_done)               # the hook reads the loss scalars
        backward_hooks.after_decoder_backward(decoder, d, allreduce is not None)
    main.wait_stream(hi)
    return d


def complete_vae_loss(encoder, decoder, property_predictor, x: torch.Tensor, conditions: torch.Tensor,
                      beta: float = 0.4, lambda_prop: float = 0.1, lambda_collapse: float = 0.01,
                      teacher_forcing_ratio: float = 0.9, free_bits: float = 0.5, lambda_mi: float = 0.0,
                      target_mi: float = 4.85, *, eps: Optional[torch.Tensor] = None, tf_mask=None, seed: int = 0,
                      pad_mask: bool = False, return_logits: bool = False) -> dict:
    """complete_vae_loss.py:7-99.  Returns the reference's 12-key dict; scalars are 0-d CUDA tensors (no host sync).

    Keyword-only extras make the stochastic inputs explicit: ``eps`` (the N(0,1) draw of reparameterize, else
    Philox(seed) on device), ``tf_mask`` (bool [T], else one ``np.random.rand()`` coin per position like
    decoder.py:180), ``pad_mask`` (opt-in masked CE; the reference is unmasked)."""
    return _run(encoder, decoder, property_predictor, x, conditions, beta, lambda_prop, lambda_collapse,
                teacher_forcing_ratio, free_bits, lambda_mi, target_mi, eps, tf_mask, seed, pad_mask, False, None,
                return_logits)


def loss_and_grad(encoder, decoder, property_predictor, x: torch.Tensor, conditions: torch.Tensor, *,
                  beta: float = 0.4, lambda_prop: float = 0.1, lambda_collapse: float = 0.01,
                  teacher_forcing_ratio: float = 0.9, free_bits: float = 0.5, lambda_mi: float = 0.0,
                  target_mi: float = 4.85, eps: Optional[torch.Tensor] = None, tf_mask=None, seed: int = 0,
                  pad_mask: bool = False, zero_grad: bool = True, allreduce=None):
    """``mx.value_and_grad(model_loss_fn, argnums=[0,1])(encoder, decoder, x, conditions)`` (trainer.py:292, :305):
    returns (loss dict, (enc_grads, dec_grads)) with the gradients as nested dicts keyed like the parameter trees
    (views of ``encoder.grads`` / ``decoder.grads``; parameters the loss does not reach hold exact zeros)."""
    if zero_grad:
        encoder.zero_grad()
        decoder.zero_grad()
    d = _run(encoder, decoder, property_predictor, x, conditions, beta, lambda_prop, lambda_collapse,
             teacher_forcing_ratio, free_bits, lambda_mi, target_mi, eps, tf_mask, seed, pad_mask, True, allreduce)
    return d, (encoder.gradients(), decoder.gradients())
