"""mutual_information / posterior_collapse — drop-in for losses/info.py of the reference."""
import torch

from ._fused import fused_loss, make_hyper


def mutual_information(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    """max(mean_b KL_b - KL(N(mean mu, mean var) || N(0,I)), 0)  (losses/info.py:3-50)."""
    return fused_loss(None, None, mu, logvar, make_hyper(), want_grads=False, want_z=False).scalar("mutual_info")


def posterior_collapse(mu: torch.Tensor, logvar: torch.Tensor, target_mi: float = 4.85, weight: float = 0.1) -> torch.Tensor:
    """weight * max(0, target_mi - MI)  (losses/info.py:53-78)."""
    hp = make_hyper(lambda_collapse=weight, collapse_target_mi=target_mi)
    return fused_loss(None, None, mu, logvar, hp, want_grads=False, want_z=False).scalar("collapse_penalty")
