"""Python face of the fused loss kernel (csrc/loss.cu, ``arcvae_loss_fwd_bwd``)."""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib


class FusedLossOut:
    __slots__ = ("losses", "dlogits", "dmu", "dlogvar", "z", "stats")

    def scalar(self, key: str) -> torch.Tensor:
        return self.losses[_lib.LOSS_KEYS.index(key)]


def make_hyper(beta=0.0, lambda_prop=0.0, lambda_collapse=0.0, free_bits=0.0, lambda_mi=0.0, target_mi=4.85,
               collapse_target_mi=4.85, pad_mask=False) -> _lib.LossHyper:
    return _lib.LossHyper(float(beta), float(lambda_prop), float(lambda_collapse), float(free_bits), float(lambda_mi),
                          float(target_mi), float(collapse_target_mi), 1 if pad_mask else 0)


def fused_loss(logits: Optional[torch.Tensor], targets: Optional[torch.Tensor], mu: Optional[torch.Tensor],
               logvar: Optional[torch.Tensor], hyper: _lib.LossHyper, *, eps: Optional[torch.Tensor] = None,
               seed: int = 0, offset: int = 0, pad_token: int = 0, want_grads: bool = True, want_z: bool = True,
               inplace_dlogits: bool = False, allreduce=None, stats: Optional[torch.Tensor] = None,
               ce_T: int = 0, bufs: Optional[dict] = None) -> FusedLossOut:
    """One launch (two around ``allreduce`` under data parallelism) computing every scalar of complete_vae_loss and,
    if ``want_grads``, d total / d logits, d mu, d logvar.  ``logits`` is [B,T,V] with unit stride on V and arbitrary
    (b,t) strides; ``targets`` [B,T] int32 with arbitrary strides.  ``allreduce(stats)`` (optional) must sum the first
    2L+5 doubles of ``stats`` across ranks in place on the current stream.
    ``stats`` + ``ce_T`` (with ``logits=None``): the cross-entropy sum was already accumulated into ``stats[2L+2]`` by
    ``decoder.forward_ce`` (fc_out epilogue); the kernel only adds the token count B*ce_T and reports recon / total.
    ``bufs`` (optional): caller-owned outputs ``losses``, ``dmu``, ``dlogvar``, ``z`` (side-stream callers allocate on
    their own stream)."""
    lib = _lib.load()
    _lib.require_cuda(logits, targets, mu, logvar, eps)
    out = FusedLossOut()
    dev = (logits if logits is not None else mu).device
    B = (logits.shape[0] if logits is not None else mu.shape[0])
    T = V = 0
    ls_b = ls_t = ts_b = ts_t = 0
    if logits is not None:
        if logits.dtype != torch.float32 or logits.stride(2) != 1:
            logits = logits.float().contiguous()
        if targets.dtype != torch.int32:
            targets = targets.to(torch.int32)
        B, T, V = logits.shape
        ls_b, ls_t = logits.stride(0), logits.stride(1)
        ts_b, ts_t = targets.stride(0), targets.stride(1)
    L = 0
    if mu is not None:
        mu = mu.float().contiguous()
        logvar = logvar.float().contiguous()
        L = mu.shape[1]
        if eps is not None:
            eps = eps.float().contiguous()
    ce_pre = 4 if (logits is None and ce_T > 0) else 0
    if ce_pre:
        T = int(ce_T)
    out.stats = stats if stats is not None else torch.zeros(2 * L + 6, dtype=torch.float64, device=dev)
    bufs = bufs or {}
    out.losses = bufs.get("losses") if bufs.get("losses") is not None else torch.empty(len(_lib.LOSS_KEYS), dtype=torch.float32, device=dev)
    out.dlogits = out.dmu = out.dlogvar = out.z = None
    if logits is not None and want_grads:
        out.dlogits = logits if inplace_dlogits else torch.empty_strided(logits.shape, logits.stride(),
                                                                       dtype=torch.float32, device=dev)
    if mu is not None:
        if want_grads:
            out.dmu = bufs["dmu"] if "dmu" in bufs else torch.empty_like(mu)
            out.dlogvar = bufs["dlogvar"] if "dlogvar" in bufs else torch.empty_like(logvar)
        if want_z:
            out.z = bufs["z"] if "z" in bufs else torch.empty_like(mu)

    def call(phases):
        _lib.check(lib.arcvae_loss_fwd_bwd(_lib.ptr(logits), ls_b, ls_t, _lib.ptr(targets), ts_b, ts_t, B, T, V,
                                           pad_token, _lib.ptr(mu), _lib.ptr(logvar), _lib.ptr(eps), L, hyper,
                                           seed, offset, phases | ce_pre, out.stats.data_ptr(), out.losses.data_ptr(),
                                           _lib.ptr(out.dlogits), _lib.ptr(out.dmu), _lib.ptr(out.dlogvar),
                                           _lib.ptr(out.z), _lib.stream_ptr()))

    if allreduce is None:
        call(3)
    else:
        call(1)
        allreduce(out.stats[:2 * L + 5])
        call(2)
    return out
