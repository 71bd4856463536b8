"""Fused train-mode BatchNorm2d + LeakyReLU + Dropout on channels-last bf16 activations
(utilities/UAPS_unet.py:37-43), over uaps_bn_* of the C ABI.  Saves only the conv output and the
batch mean / rstd for backward; the LeakyReLU sign and the dropout mask are recomputed."""
from __future__ import annotations

import torch

from . import _lib as L
from . import stepctx
from .perturb import _next_seed


class _BnActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, momentum, eps, slope, p_drop, seed, seed_dev, sums_pre, nrep):
        L.require_cuda(y)
        if y.dtype != torch.bfloat16 or not y.is_contiguous(memory_format=torch.channels_last):
            raise RuntimeError("bn_act expects a channels_last bf16 [B,C,H,W] tensor")
        B, C, H, W = y.shape
        npix = B * H * W
        dev = y.device
        sc = stepctx.current()
        if sums_pre is not None:                   # the producing conv's epilogue already accumulated them (nrep replicas)
            sums = sums_pre
        else:
            nrep = 1
            sums = sc.take(2 * C) if sc is not None else torch.zeros(2 * C, dtype=torch.float64, device=dev)
        stats = torch.empty(2 * C, dtype=torch.float32, device=dev)          # save_mean | save_rstd
        out = torch.empty_like(y)
        lib = L.lib()
        g32 = gamma.data if gamma.dtype == torch.float32 and gamma.is_contiguous() else gamma.detach().float().contiguous()
        b32 = beta.data if beta.dtype == torch.float32 and beta.is_contiguous() else beta.detach().float().contiguous()
        with L.on_device(dev):
            if sums_pre is None:
                L.check(lib.uaps_bn_stats_nhwc(y.data_ptr(), npix, C, sums.data_ptr(), sums[C:].data_ptr(), L.stream_ptr()),
                        "uaps_bn_stats_nhwc")
            L.check(lib.uaps_bn_act_nhwc(y.data_ptr(), sums.data_ptr(), sums[C:].data_ptr(), g32.data_ptr(), b32.data_ptr(),
                                         None if running_mean is None else running_mean.data_ptr(),
                                         None if running_var is None else running_var.data_ptr(),
                                         momentum, eps, slope, p_drop, seed, out.data_ptr(), stats.data_ptr(),
                                         stats[C:].data_ptr(), npix, C, seed_dev, int(nrep), L.stream_ptr()), "uaps_bn_act_nhwc")
        ctx.save_for_backward(y, g32, b32, stats)
        ctx.cfg = (slope, p_drop, seed, seed_dev)
        ctx.params = (gamma, beta)
        return out

    @staticmethod
    def backward(ctx, g):
        y, g32, b32, stats = ctx.saved_tensors
        slope, p_drop, seed, seed_dev = ctx.cfg
        B, C, H, W = y.shape
        g = g.contiguous(memory_format=torch.channels_last)
        if g.dtype != torch.bfloat16:
            g = g.to(torch.bfloat16)
        sc = stepctx.current()
        sums = sc.take(2 * C) if sc is not None else torch.zeros(2 * C, dtype=torch.float64, device=y.device)
        dy = torch.empty_like(y)
        gamma, beta = ctx.params
        direct = (sc is not None and sc.direct_grads and gamma.grad is not None and beta.grad is not None
                  and gamma.grad.dtype == torch.float32 and gamma.grad.is_contiguous() and beta.grad.is_contiguous())
        with L.on_device(y.device):
            L.check(L.lib().uaps_bn_act_bwd_nhwc(g.data_ptr(), y.data_ptr(), g32.data_ptr(), b32.data_ptr(), stats.data_ptr(),
                                                 stats[C:].data_ptr(), slope, p_drop, seed, sums.data_ptr(),
                                                 sums[C:].data_ptr(), dy.data_ptr(),
                                                 gamma.grad.data_ptr() if direct else None,
                                                 beta.grad.data_ptr() if direct else None,
                                                 B * H * W, C, seed_dev, L.stream_ptr()), "uaps_bn_act_bwd_nhwc")
        if direct:                               # already added into gamma.grad / beta.grad by the kernel
            return (dy,) + (None,) * 12
        # the kernel accumulates sum(g') and the RAW sum(g' * y); d gamma = sum(g' * xhat) = rstd * (sum(g' y) - mean * sum(g'))
        sg, sgy = sums[:C], sums[C:]
        dgamma = stats[C:].double() * (sgy - stats[:C].double() * sg)
        return (dy, dgamma.float(), sg.float()) + (None,) * 10


def bn_lrelu_dropout(y: torch.Tensor, bn: torch.nn.BatchNorm2d, p_drop: float = 0.0, slope: float = 0.01,
                     seed=None, sums: torch.Tensor = None, nrep: int = 1) -> torch.Tensor:
    """dropout(leaky_relu(batch_norm(y))) with batch statistics; advances bn's running statistics like
    nn.BatchNorm2d in training mode.  sums / nrep: batch sums already accumulated by the conv that produced y
    (``conv_bf16(..., bn_sums=, bn_nrep=)``): the statistics pass over y is skipped."""
    seed_dev = None
    if p_drop > 0.0 and seed is None:
        sc = stepctx.current()
        if sc is not None and sc.state is not None:       # device-resident step: per-call constant + the iteration's key
            seed, seed_dev = sc.next_seed(), sc.state.ptr("key_rank")
        else:
            seed = _next_seed()
    out = _BnActFn.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.momentum), float(bn.eps),
                         float(slope), float(p_drop), 0 if seed is None else int(seed), seed_dev, sums, int(nrep))
    if bn.num_batches_tracked is not None:
        sc = stepctx.current()
        if sc is not None:
            sc.counters.append(bn.num_batches_tracked)      # one foreach add at the end of the iteration
        else:
            bn.num_batches_tracked += 1
    return out
