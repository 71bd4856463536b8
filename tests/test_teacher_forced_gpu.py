"""GPU: TEACHER-FORCED per-layer parity of the bf16 / tcgen05 network against the oracle (north_star: "1e-2 for bf16
conv activations").

A 22-layer bf16 network is chaotic in its roundings (a 1-ulp flip moves a LeakyReLU / BatchNorm input), so comparing whole
networks says little.  Here every hand-written layer of the product -- the 124 convolutions / BatchNorm+LeakyReLU / pool /
upsample steps of one forward of the K = 4 model -- is fed the ORACLE's own input activation of that layer
(oracle/unet_ref.py:unet_layer_trace, the fp32 restatement that oracle/make_golden.py pins to the reference's UNet_UAPS)
and must reproduce the oracle's output of that layer to 1e-2 of its scale.  Backward: each layer is fed the oracle's own
upstream gradient and compared with fp32 torch autograd of the same layer on the same (bf16-rounded) operands: input
gradients, weight / bias gradients, BatchNorm gamma / beta gradients.

Cases: the reference-generated golden input (tests/golden/unet_small.npz), 2x3x256x256 (NEU shape), 1x3x240x640 with C = 2
(KoSDD2 shape), 1x1x512x512 with C = 2 (DAGM shape).  Weights: fan-in scaled normal (He) convolutions, BN gamma ~ 1.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle.unet_ref import feature_shapes, synthetic_rand, synthetic_state_dict, unet_layer_trace

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _cl16(t):
    """fp32 NCHW -> channels-last bf16 (logical NCHW) as the bf16 path stores activations."""
    from uaps_b200.conv import to_nhwc_bf16
    if t.shape[1] % 16:
        return to_nhwc_bf16(t).permute(0, 3, 1, 2)                      # the 1- / 3-channel network input: 16-padded
    return t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def _r(t):
    return t.detach().to(torch.bfloat16).float()


def _close(got, want, what, tol=TOL):
    got, want = got.float(), want.float()
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= tol * scale + 1e-6, f"{what}: max err {err:.3e} vs scale {scale:.3e} ({err / max(scale, 1e-30):.3e} rel)"
    return err / max(scale, 1e-30)


def _case(name):
    dev = torch.device("cuda:0")
    if name == "golden":
        x = torch.from_numpy(np.load(os.path.join(GOLDEN, "unet_small.npz"))["x"])
        cin, C, seed = 3, 4, 1234
    elif name == "neu256":
        x, cin, C, seed = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(0)), 3, 4, 7
    elif name == "kosdd2":
        x, cin, C, seed = torch.randn(1, 3, 240, 640, generator=torch.Generator().manual_seed(1)), 3, 2, 8
    else:
        x, cin, C, seed = torch.randn(1, 1, 512, 512, generator=torch.Generator().manual_seed(2)), 1, 2, 9
    sd = {k: v.to(dev) for k, v in synthetic_state_dict(cin, C, seed=seed).items()}
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    B, _, H, W = x.shape
    rand = synthetic_rand(feature_shapes(B, H, W), seed=5)
    rand = {k: [v.to(dev) if torch.is_tensor(v) else v for v in vals] for k, vals in rand.items()}
    return x.to(dev).requires_grad_(True), sd, rand


@pytest.mark.parametrize("case", ["golden", "neu256", "kosdd2", "dagm512"])
def test_every_layer_teacher_forced(case):
    from uaps_b200.bn_act import bn_lrelu_dropout
    from uaps_b200.conv import conv_bf16
    from uaps_b200.resample import maxpool2, upsample2x
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x, sd, rand = _case(case)
    outs, trace = unet_layer_trace(x, sd, rand)
    g = torch.Generator(device=x.device).manual_seed(3)
    sum((o * torch.randn(o.shape, generator=g, device=o.device)).sum() for o in outs).backward()       # oracle upstream gradients
    worst = {}
    for rec in trace:
        kind, name, ref_out, g_out = rec["kind"], rec["name"], rec["out"].detach(), rec["out"].grad
        assert g_out is not None, name
        ins = [t.detach() for t in rec["inputs"]]
        if kind in ("conv", "logits"):
            w, b = sd[rec["weight"]].detach(), sd[rec["bias"]].detach()
            wp, bp = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
            xs = [_cl16(t).requires_grad_(True) for t in ins]
            y = conv_bf16(xs[0], wp, bp, x2=xs[1] if len(xs) > 1 else None, nchw_f32_out=(kind == "logits"),
                          bias_grad=not rec["before_bn"])
            e = _close(y, ref_out, f"{case}/{name} fwd")
            # backward: the oracle's upstream gradient through the custom dgrad / wgrad kernels vs fp32 autograd of the same conv
            gy = g_out if kind == "logits" else g_out.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            y.backward(gy)
            xr = [_r(t).requires_grad_(True) for t in ins]
            wr, br = _r(w).requires_grad_(True), b.clone().requires_grad_(True)
            yr = F.conv2d(xr[0] if len(xr) == 1 else torch.cat(xr, 1), wr, br, padding=w.shape[-1] // 2)
            grads = torch.autograd.grad(yr, xr + [wr, br], g_out if kind == "logits" else _r(g_out))
            for i, (xc, gr) in enumerate(zip(xs, grads[:len(xs)])):
                if name == "encoder.in_conv.conv_conv.0":
                    continue                                     # no gradient is needed for the image itself
                _close(xc.grad[:, :gr.shape[1]], gr, f"{case}/{name} dgrad[{i}]")
            _close(wp.grad, grads[len(xs)], f"{case}/{name} wgrad")
            if not rec["before_bn"]:
                _close(bp.grad, grads[len(xs) + 1], f"{case}/{name} bias grad", tol=2e-2)
        elif kind == "bn_act":
            bn = torch.nn.BatchNorm2d(ref_out.shape[1]).to(x.device).train()
            with torch.no_grad():
                bn.weight.copy_(sd[rec["bn"] + ".weight"]); bn.bias.copy_(sd[rec["bn"] + ".bias"])
            yc = _cl16(ins[0]).requires_grad_(True)
            a = bn_lrelu_dropout(yc, bn, 0.0)
            e = _close(a, ref_out, f"{case}/{name} fwd")
            a.backward(g_out.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
            yr = _r(ins[0]).requires_grad_(True)
            gam, bet = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
            ar = F.leaky_relu(F.batch_norm(yr, None, None, gam, bet, True, 0.1, 1e-5), 0.01)
            gyr, ggam, gbet = torch.autograd.grad(ar, [yr, gam, bet], _r(g_out))
            _close(yc.grad, gyr, f"{case}/{name} dy", tol=2e-2)            # one bf16 rounding of dy itself is 4e-3 of the value
            _close(bn.weight.grad, ggam, f"{case}/{name} dgamma")
            _close(bn.bias.grad, gbet, f"{case}/{name} dbeta")
        elif kind in ("maxpool", "upsample"):
            fn = maxpool2 if kind == "maxpool" else upsample2x
            xc = _cl16(ins[0]).requires_grad_(True)
            if ins[0].shape[1] % 16 == 0:                                 # (the padded 3-channel image is never pooled)
                yv = fn(xc)
                e = _close(yv, ref_out, f"{case}/{name} fwd")
                yv.backward(g_out.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
                xr = _r(ins[0]).requires_grad_(True)
                yr = F.max_pool2d(xr, 2) if kind == "maxpool" else F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
                (gxr,) = torch.autograd.grad(yr, [xr], _r(g_out))
                _close(xc.grad, gxr, f"{case}/{name} dx", tol=2e-2)
            else:
                e = 0.0
        else:
            raise AssertionError(kind)
        worst[kind] = max(worst.get(kind, 0.0), e)
    print(f"{case}: {len(trace)} layers, worst forward error / scale by kind: " + ", ".join(f"{k} {v:.2e}" for k, v in sorted(worst.items())))
    assert len(trace) == 124
