"""Oracle: the loss section of UAPS_train.py restated as functions (TEST INFRASTRUCTURE).

The reference has this code inline in its training loop (no function to
import), written out for exactly K = 4 decoders.  Here the same expressions
are applied in the same order (left-associated sums, the same torch modules)
for a list of K logits tensors.  Reference line numbers are for
``/root/reference/UAPS_train.py`` unless another file is named.

Used by tests/, smoke() and bench.py's CPU baseline only.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# UAPS_train.py:73-75 -- the three module-level criteria of the reference.
_kl_none = nn.KLDivLoss(reduction="none")
_log_sm = nn.LogSoftmax(dim=1)
_ce = nn.CrossEntropyLoss()


def dice_loss_ref(true: torch.Tensor, logits: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """utilities/pytorch_losses.py:54-89 (multi-class branch :81-89).

    One declared deviation: the identity matrix is created on ``true``'s
    device.  The reference indexes a CPU ``torch.eye`` with the label tensor
    (:81), which raises on torch >= 2 when the labels live on CUDA; on CPU the
    two are the same expression.
    """
    c = logits.shape[1]
    if c == 1:
        raise NotImplementedError("binary (C == 1) branch is not on the UAPS path")
    one_hot = torch.eye(c, device=true.device)[true.squeeze(1)]          # :81
    one_hot = one_hot.permute(0, 3, 1, 2).float()                        # :82
    probas = F.softmax(logits, dim=1)                                    # :83
    one_hot = one_hot.type(logits.type())                                # :84
    dims = (0,) + tuple(range(2, true.ndimension()))                     # :85
    inter = torch.sum(probas * one_hot, dims)                            # :86
    card = torch.sum(probas + one_hot, dims)                             # :87
    return 1 - (2.0 * inter / (card + eps)).mean()                       # :88-89


def sigmoid_rampup_ref(current: float, rampup_length: float) -> float:
    """utilities/ramps.py:19-26."""
    if rampup_length == 0:
        return 1.0
    cur = float(np.clip(current, 0.0, rampup_length))
    phase = 1.0 - cur / rampup_length
    return float(math.exp(-5.0 * phase * phase))


def consistency_weight_ref(iter_num: int, consistency: float = 0.1, rampup: float = 200.0,
                           iters_per_ramp_epoch: int = 80) -> float:
    """UAPS_train.py:81-87 and :279-280 (``iter_num // 80``)."""
    return consistency * sigmoid_rampup_ref(iter_num // iters_per_ramp_epoch, rampup)


def unlabeled_loss_ref(logits: Sequence[torch.Tensor], mix_w: Sequence[float],
                       cw1: float, cw2: float):
    """UAPS_train.py:186-189, 223-282 for K = len(logits) decoders.

    ``mix_w`` is the injected Dirichlet draw of :251 (float64 numpy scalars in
    the reference; python floats here -- both are rounded to fp32 by the
    scalar * tensor multiply).

    Returns a dict with the differentiable scalars and the per-pixel maps.
    """
    K = len(logits)
    soft = [torch.softmax(z, dim=1) for z in logits]                     # :186-189
    acc = soft[0]
    for k in range(1, K):
        acc = acc + soft[k]
    preds = acc / K                                                      # :223
    var = [torch.sum(_kl_none(_log_sm(z), preds), dim=1) for z in logits]  # :226,229,232,235
    exp_var = [torch.exp(-v) for v in var]                               # :227,230,233,236
    vsum = var[0]
    for k in range(1, K):
        vsum = vsum + var[k]
    ave_var = vsum / K                                                   # :241
    l_uncert = torch.mean(ave_var)                                       # :243

    mixed = float(mix_w[0]) * soft[0].detach()                           # :252-255
    for k in range(1, K):
        mixed = mixed + float(mix_w[k]) * soft[k].detach()
    pseudo = torch.argmax(mixed, dim=1, keepdim=False)

    ps = [0.5 * (_ce(z, pseudo) + dice_loss_ref(pseudo.unsqueeze(1), z)) for z in logits]  # :259-262
    ps_w = [torch.mean(ps[k] * exp_var[k]) for k in range(K)]            # :265-268
    tot = ps_w[0]
    for k in range(1, K):
        tot = tot + ps_w[k]
    ps_loss = tot / K                                                    # :277
    loss_u = cw1 * ps_loss + cw2 * l_uncert                              # :282 (unlabeled terms)
    return {"loss_u": loss_u, "ps_loss": ps_loss, "l_uncert": l_uncert, "pseudo": pseudo,
            "var": var, "exp_var": exp_var, "ps": ps, "soft": soft, "mixed": mixed}


def supervised_loss_ref(logits: Sequence[torch.Tensor], labels: torch.Tensor):
    """UAPS_train.py:194-218: mean over decoders of 0.5 * (CE + Dice) against ground truth."""
    K = len(logits)
    ce = [_ce(z, labels.long()) for z in logits]                         # :194-197
    dice = [dice_loss_ref(labels.unsqueeze(1), z) for z in logits]       # :201-204
    each = [0.5 * (ce[k] + dice[k]) for k in range(K)]                   # :208-211
    sup, tce, tdice = each[0], ce[0], dice[0]
    for k in range(1, K):
        sup, tce, tdice = sup + each[k], tce + ce[k], tdice + dice[k]
    return {"supervised_loss": sup / K, "total_loss_ce": tce / K,       # :216-218
            "total_loss_dice": tdice / K, "ce": ce, "dice": dice}


def unlabeled_loss_fp64_closed_form(logits: Sequence[torch.Tensor], mix_w, lam1: float, lam2: float,
                                    pseudo: torch.Tensor):
    """fp64 closed form of d(lam1*ps_loss + lam2*l_uncert)/d(logits).

    Independent of autograd; used to cross-check both the autograd of
    ``unlabeled_loss_ref`` and the CUDA pass-2 kernel.  ``pseudo`` is given
    (argmax is not differentiable and is decided in fp32 by the caller).
    """
    K = len(logits)
    z = [t.detach().double() for t in logits]
    B, C, H, W = z[0].shape
    N = B * H * W
    p = [torch.softmax(t, 1) for t in z]
    l = [torch.log_softmax(t, 1) for t in z]
    q = sum(p) / K
    xlogx = torch.where(q > 0, q * torch.log(q), torch.zeros_like(q))
    V = [(xlogx - q * l[k]).sum(1) for k in range(K)]
    E = [torch.exp(-v) for v in V]
    onehot = F.one_hot(pseudo, C).permute(0, 3, 1, 2).double()
    eps = 1e-7
    ps, Ebar, consts = [], [], []
    for k in range(K):
        ce = -(l[k] * onehot).sum() / N
        I = (p[k] * onehot).sum((0, 2, 3))
        card = (p[k] + onehot).sum((0, 2, 3))
        dice = 1 - (2 * I / (card + eps)).mean()
        ps.append(0.5 * (ce + dice))
        Ebar.append(E[k].mean())
        consts.append((I, card))
    ps_loss = sum(ps[k] * Ebar[k] for k in range(K)) / K
    l_unc = sum(v.mean() for v in V) / K
    g = [(lam2 - lam1 * ps[k] * E[k]) / (N * K) for k in range(K)]       # [B,H,W]
    Gq = sum(g[k].unsqueeze(1) * (torch.log(q) + 1 - l[k]) for k in range(K))
    dz = []
    for k in range(K):
        I, card = consts[k]
        Gl = -g[k].unsqueeze(1) * q - lam1 * Ebar[k] / (2 * K * N) * onehot
        den = (card + eps).view(1, C, 1, 1)
        Gp = Gq / K - lam1 * Ebar[k] / (2 * K) * (2.0 / C) * (onehot * den - I.view(1, C, 1, 1)) / den ** 2
        d = Gl - p[k] * Gl.sum(1, keepdim=True) + p[k] * (Gp - (p[k] * Gp).sum(1, keepdim=True))
        dz.append(d)
    return {"ps_loss": ps_loss, "l_uncert": l_unc, "dz": dz, "ps": ps, "Ebar": Ebar}
