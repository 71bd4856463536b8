"""Bilinear x2 upsampling (align_corners=True) and 2x2 max pooling on channels-last bf16 activations
(utilities/UAPS_unet.py:74-75 and :56), over uaps_upsample2x_nhwc / uaps_maxpool2_nhwc."""
from __future__ import annotations

import torch

from . import _lib as L


def _check_cl(x: torch.Tensor) -> None:
    L.require_cuda(x)
    if x.dtype != torch.bfloat16 or not x.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError("expected a channels_last bf16 [B,C,H,W] tensor")


def _empty_cl(B, C, H, W, dev):
    return torch.empty((B, C, H, W), dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last) \
        if False else torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dev).permute(0, 3, 1, 2)


class _Upsample2xFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _check_cl(x)
        B, C, H, W = x.shape
        y = _empty_cl(B, C, 2 * H, 2 * W, x.device)
        with L.on_device(x.device):
            L.check(L.lib().uaps_upsample2x_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, C, 0, L.stream_ptr()), "uaps_upsample2x_nhwc")
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        g = g.contiguous(memory_format=torch.channels_last)
        gx = _empty_cl(B, C, H, W, g.device)
        with L.on_device(g.device):
            L.check(L.lib().uaps_upsample2x_nhwc(g.data_ptr(), gx.data_ptr(), B, H, W, C, 1, L.stream_ptr()),
                    "uaps_upsample2x_nhwc(bwd)")
        return gx


class _MaxPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _check_cl(x)
        B, C, H, W = x.shape
        y = _empty_cl(B, C, H // 2, W // 2, x.device)
        with L.on_device(x.device):
            L.check(L.lib().uaps_maxpool2_nhwc(x.data_ptr(), None, y.data_ptr(), B, H, W, C, L.stream_ptr()), "uaps_maxpool2_nhwc")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        B, C, H, W = x.shape
        g = g.contiguous(memory_format=torch.channels_last)
        gx = torch.empty_like(x)
        with L.on_device(x.device):
            L.check(L.lib().uaps_maxpool2_nhwc(x.data_ptr(), g.data_ptr(), gx.data_ptr(), B, H, W, C, L.stream_ptr()),
                    "uaps_maxpool2_nhwc(bwd)")
        return gx


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    return _Upsample2xFn.apply(x)


def maxpool2(x: torch.Tensor) -> torch.Tensor:
    return _MaxPool2Fn.apply(x)
