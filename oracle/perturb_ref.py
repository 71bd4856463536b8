"""Oracle: the three encoder-feature perturbations with the randomness injected
(TEST INFRASTRUCTURE).  Line numbers: /root/reference/utilities/UAPS_unet.py.

The reference draws its randomness from three different global generators
(torch CPU RNG :178, torch device RNG inside F.dropout :157, numpy global
:165), so parity is defined with the random draws passed in.
"""
from __future__ import annotations

import torch


def feature_noise_ref(x: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """:172-185.  ``noise`` has shape x.shape[1:] (one draw of U(-r, r) shared by the batch, :178-179)."""
    assert noise.shape == x.shape[1:]
    return x.mul(noise.to(x.device).unsqueeze(0)) + x                    # :179-180


def dropout_ref(x: torch.Tensor, keep_mask: torch.Tensor, p: float = 0.5) -> torch.Tensor:
    """:156-158.  F.dropout(x, p) with training=True: kept elements scaled by 1/(1-p)."""
    return x * keep_mask.to(x.dtype) * (1.0 / (1.0 - p))


def feature_dropout_ref(x: torch.Tensor, u: float) -> torch.Tensor:
    """:161-169 with the single ``np.random.uniform(0.7, 0.9)`` draw of :165 passed as ``u``."""
    attention = torch.mean(x, dim=1, keepdim=True)                       # :162
    max_val, _ = torch.max(attention.view(x.size(0), -1), dim=1, keepdim=True)  # :163-164
    threshold = max_val * u                                              # :165
    threshold = threshold.view(x.size(0), 1, 1, 1).expand_as(attention)  # :166
    drop_mask = (attention < threshold).float()                          # :167
    return x.mul(drop_mask)                                              # :168
