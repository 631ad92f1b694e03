"""Mirrors losses/__init__.py of the reference (the four loss modules the training path uses)."""
from .info import mutual_information, posterior_collapse
from .kl import kl_divergence
from .prop import property_prediction_loss
from .recon import reconstruction_loss

__all__ = ["reconstruction_loss", "kl_divergence", "mutual_information", "posterior_collapse",
           "property_prediction_loss"]
