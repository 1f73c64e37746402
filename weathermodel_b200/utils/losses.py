"""Loss terms with the reference's names (src/utils/losses.py:10-87). These torch-op versions serve
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


def compute_mixture_kl_divergence(z, feature_mask, mu_x, var_x, mu_k, var_k, log_w_k):
    """Single-sample KL(q(z|x) || sum_i w_i N(mu_k[i], var_k[i])) = log q(z|x) - log p(z) over masked entries
    (reference src/utils/losses.py:50-87). z, mu_x, var_x, feature_mask [B,S,F]; mu_k, var_k [B,k,S,F]; log_w_k [B,k]."""
    log_q = gaussian_log_likelihood(z, mu_x, var_x, feature_mask, (1, 2))
    log_comp = gaussian_log_likelihood(z.unsqueeze(1), mu_k, var_k, feature_mask.unsqueeze(1), (2, 3))  # [B, k]
    return log_q - torch.logsumexp(log_w_k + log_comp, dim=1)
