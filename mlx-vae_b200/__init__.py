"""mlx_vae_b200 — B200-native (sm_100a) AR-CVAE training step and sampler behind the call surface of
Raiden-Makoto/MLX-VAE (models/{encoder,decoder,decoder_sampling,vae}.py, losses/{recon,kl,info,prop}.py,
complete_vae_loss.py, the step of trainer.py).  All computation happens in libarcvae_sm100.so (csrc/, C ABI in
include/arcvae_b200.h); there is no CPU or PyTorch fallback."""
from . import _lib
from .complete_vae_loss import complete_vae_loss, loss_and_grad
from .losses import (kl_divergence, mutual_information, posterior_collapse, property_prediction_loss,
                     reconstruction_loss)
from .models import ARCVAE, MLXAutoregressiveDecoder, MLXAutoregressiveDecoderSampling, MLXEncoder
from .trainer import ARCVAETrainerWithLoss
from .data import MoleculeDataset

__all__ = ["ARCVAE", "MLXEncoder", "MLXAutoregressiveDecoder", "MLXAutoregressiveDecoderSampling",
           "complete_vae_loss", "loss_and_grad", "reconstruction_loss", "kl_divergence", "mutual_information",
           "posterior_collapse", "property_prediction_loss", "ARCVAETrainerWithLoss", "MoleculeDataset"]
