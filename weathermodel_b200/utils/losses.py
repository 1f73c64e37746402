"""Loss terms with the reference's names (src/utils/losses.py:10-47). These torch-op versions serve
callers that hold (mu, var) tensors -- e.g. the yield fine-tune heads; the pretraining trainers use the
fused CUDA loss heads in weathermodel_b200.engine instead."""
from typing import Optional, Tuple

import torch


def gaussian_log_likelihood(x, mu, var, feature_mask, masked_dims: Optional[Tuple[int, ...]] = None):
    dims = tuple(range(1, x.ndim)) if masked_dims is None else masked_dims
    ll = -0.5 * torch.log(2 * torch.pi * var) - 0.5 * (x - mu) ** 2 / var
    return torch.sum(ll * feature_mask, dim=dims)


def compute_gaussian_kl_divergence(feature_mask, mu_x, var_x, mu_p, var_p):
    per_dim = 0.5 * (torch.log(var_p / var_x) + var_x / var_p + (mu_x - mu_p) ** 2 / var_p - 1.0)
    return torch.sum(per_dim * feature_mask, dim=(1, 2))
