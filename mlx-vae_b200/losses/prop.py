"""property_prediction_loss — mirrors losses/prop.py of the reference.

Dead on the training path: ``property_predictor`` is always None (train.py:186) and the one call site passes a
wrong argument list (complete_vae_loss.py:63-67 vs prop.py:5-11), so ``prop_loss`` is the constant 0.
Kept for API completeness only; a handful of torch element-wise ops, never launched by the step."""
import torch


def property_prediction_loss(z, predicted_properties, target_properties, property_scales=None, reduction: str = "mean"):
    mse = torch.square(predicted_properties - target_properties)          # prop.py:29
    if property_scales is not None:
        mse = mse / (torch.square(property_scales) + 1e-8)                # prop.py:31-33
    if reduction == "mean":
        return mse.mean()
    if reduction == "sum":
        return mse.sum()
    return mse
