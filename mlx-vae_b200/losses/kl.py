"""kl_divergence — drop-in for losses/kl.py of the reference."""
import torch

from ._fused import fused_loss, make_hyper


def kl_divergence(mu: torch.Tensor, logvar: torch.Tensor, reduction: str = "mean", free_bits: float = 0.0) -> torch.Tensor:
    """KL(q(z|x,c) || N(0,I)) with the reference's clips, max(.,0) and per-dimension free-bits floor
    (losses/kl.py:5-66)."""
    if reduction not in ("mean", "sum"):
        raise NotImplementedError("per-sample KL (reduction='none') is not on the training path")
    out = fused_loss(None, None, mu, logvar, make_hyper(free_bits=free_bits), want_grads=False, want_z=False)
    val = out.scalar("kl_loss")
    if reduction == "sum":
        val = val * float(mu.shape[0])
    return val
