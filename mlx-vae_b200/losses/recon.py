"""reconstruction_loss — drop-in for losses/recon.py of the reference."""
import torch

from ._fused import fused_loss, make_hyper


def reconstruction_loss(logits: torch.Tensor, targets: torch.Tensor, reduction: str = "mean", *,
                        pad_mask: bool = False, pad_token: int = 0) -> torch.Tensor:
    """Cross-entropy over all B*T positions (losses/recon.py:6-64); the reference applies NO padding mask
    (``pad_mask=True`` is the opt-in of SURVEY.md F3).  'mean' and 'sum' run in the fused kernel."""
    if reduction not in ("mean", "sum"):
        raise NotImplementedError("per-position CE (reduction='none') is not on the training path")
    out = fused_loss(logits, targets, None, None, make_hyper(pad_mask=pad_mask), pad_token=pad_token,
                     want_grads=False, want_z=False)
    val = out.scalar("recon_loss")
    if reduction == "sum":
        val = val * out.stats[3].to(torch.float32)   # mean * token count
    return val
