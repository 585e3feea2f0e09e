"""keras.losses stand-in (TEST INFRASTRUCTURE ONLY)."""
import torch


def binary_crossentropy(y_true, y_pred):
    """Keras 2.2 TF backend: clip p to [eps, 1-eps] (eps=1e-7), BCE, mean over last axis."""
    eps = 1e-7
    p = torch.clamp(y_pred, eps, 1.0 - eps)
    return -(y_true * torch.log(p) + (1.0 - y_true) * torch.log(1.0 - p)).mean(dim=-1)
