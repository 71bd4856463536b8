"""GPU parity of the perturbation kernels (UAPS_unet.py:156-185) through the C ABI: golden vectors made
from the reference's own functions, oracle on device, Philox-mode statistical and replay properties."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.perturb_ref import dropout_ref, feature_dropout_ref, feature_noise_ref

pytestmark = pytest.mark.gpu


def _g():
    return {k: v for k, v in np.load(os.path.join(GOLDEN, "perturb.npz")).items()}


def test_golden_injected():
    from uaps_b200 import perturb as P
    dev = torch.device("cuda:0")
    g = _g()
    x = torch.from_numpy(g["x"]).to(dev)
    noise, keep, u = torch.from_numpy(g["noise"]).to(dev), torch.from_numpy(g["keep"]).to(dev), float(g["u"])
    assert torch.equal(P.FeatureNoise()(x, noise=noise).cpu(), torch.from_numpy(g["y_noise"]))
    assert torch.equal(P.Dropout(x, 0.5, keep=keep).cpu(), torch.from_numpy(g["y_drop"]))
    assert torch.equal(P.FeatureDropout(x, u=u).cpu(), torch.from_numpy(g["y_fd"]))
    yn, yd, yf = P.perturb3(x, noise=noise, keep=keep, u=u)
    assert torch.equal(yn.cpu(), torch.from_numpy(g["y_noise"]))
    assert torch.equal(yd.cpu(), torch.from_numpy(g["y_drop"]))
    assert torch.equal(yf.cpu(), torch.from_numpy(g["y_fd"]))


@pytest.mark.parametrize("shape", [(4, 16, 256, 256), (2, 256, 16, 16), (3, 32, 25, 29), (2, 64, 58, 160)])
def test_against_oracle_with_gradients(shape):
    from uaps_b200 import perturb as P
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=gen).to(dev)
    noise = ((torch.rand(shape[1:], generator=gen) * 2 - 1) * 0.3).to(dev)
    keep = (torch.rand(shape, generator=gen) >= 0.5).to(dev)
    u = 0.8123
    cots = [torch.randn(shape, generator=gen).to(dev) for _ in range(3)]

    xr = x.clone().requires_grad_(True)
    refs = [feature_noise_ref(xr, noise), dropout_ref(xr, keep, 0.5), feature_dropout_ref(xr, u)]
    sum((r * c).sum() for r, c in zip(refs, cots)).backward()

    xo = x.clone().requires_grad_(True)
    outs = P.perturb3(xo, noise=noise, keep=keep, u=u)
    sum((o * c).sum() for o, c in zip(outs, cots)).backward()
    assert torch.equal(outs[0], refs[0]) and torch.equal(outs[1], refs[1])
    # the drop mask compares a channel-mean against a threshold: summation order may flip a pixel that
    # sits within an ulp of it; allow at most a handful of such pixels
    bad = (outs[2] != refs[2]).flatten(2).any(1).sum().item() if outs[2].dim() == 4 else 0
    assert bad <= 2, bad
    if bad == 0:
        assert torch.allclose(xo.grad, xr.grad, rtol=1e-6, atol=1e-6)

    # separate entry points too
    xs = x.clone().requires_grad_(True)
    ys = [P.FeatureNoise()(xs, noise=noise), P.Dropout(xs, 0.5, keep=keep), P.FeatureDropout(xs, u=u)]
    sum((o * c).sum() for o, c in zip(ys, cots)).backward()
    assert torch.equal(ys[0], refs[0]) and torch.equal(ys[1], refs[1])
    if bad == 0:
        assert torch.equal(ys[2], refs[2])
        assert torch.allclose(xs.grad, xr.grad, rtol=1e-6, atol=1e-6)


def test_encoder_dropout_rates():
    """nn.Dropout(p) of the encoder blocks (:40), p = 0.05 .. 0.5, injected masks vs torch."""
    from uaps_b200 import perturb as P
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(2, 16, 64, 64, generator=gen).to(dev)
    for p in (0.05, 0.1, 0.2, 0.3, 0.5):
        keep = (torch.rand(x.shape, generator=gen) >= p).to(dev)
        ref = dropout_ref(x, keep, p)
        out = P.Dropout(x, p, keep=keep)
        assert torch.allclose(out, ref, rtol=2e-7, atol=0)


def test_philox_mode_statistics_and_replay():
    from uaps_b200 import perturb as P
    dev = torch.device("cuda:0")
    x = torch.ones(8, 32, 64, 64, device=dev, requires_grad=True)
    yn, yd, yf = P.perturb3(x, seed=1234, u=0.8)
    n = yn.detach() - 1.0                                   # x = 1 -> y = 1 + noise
    assert torch.equal(n[0], n[5])                          # one draw shared by the batch
    assert n.abs().max().item() <= 0.3 and abs(n.mean().item()) < 5e-3
    assert n.std().item() == pytest.approx(0.3 / 3 ** 0.5, rel=0.02)
    frac = (yd.detach() != 0).float().mean().item()
    assert frac == pytest.approx(0.5, abs=5e-3)
    assert set(yd.detach().unique().tolist()) <= {0.0, 2.0}
    assert not torch.equal(yd.detach()[0], yd.detach()[1])  # dropout differs per sample
    # backward regenerates the same draw from the seed (no stored mask)
    (yn.sum() + yd.sum()).backward()
    assert torch.allclose(x.grad, (1.0 + n) + yd.detach(), rtol=1e-6, atol=1e-6)
    # same seed -> same draw; other seed -> other draw
    yn2, yd2, _ = P.perturb3(x.detach(), seed=1234, u=0.8)
    yn3, _, _ = P.perturb3(x.detach(), seed=99, u=0.8)
    assert torch.equal(yn2, yn.detach()) and torch.equal(yd2, yd.detach()) and not torch.equal(yn3, yn2)
    # stand-alone entry points draw the same stream as the fused kernel
    assert torch.equal(P.FeatureNoise()(x.detach(), seed=1234), yn.detach())
    assert torch.equal(P.Dropout(x.detach(), 0.5, seed=1234), yd.detach())


@pytest.mark.parametrize("shape", [(4, 16, 64, 64), (2, 64, 32, 32), (2, 256, 16, 16), (2, 128, 15, 40)])
def test_channels_last_bf16_kernel(shape):
    """bf16 NHWC perturb3 (Philox): semantics against torch expressions built from its own outputs."""
    from uaps_b200 import perturb as P
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g).abs() + 0.5).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    u = 0.8
    yn, yd, yf = P.perturb3_nhwc(x, seed=42, u=u)
    assert yn.is_contiguous(memory_format=torch.channels_last) and yn.dtype == torch.bfloat16
    xf = x.detach().float()
    # FeatureNoise: y = x (1 + n), |n| <= 0.3, one n per (c,h,w) shared by the batch
    n = yn.float() / xf - 1.0
    assert n.abs().max().item() <= 0.3 + 2e-2
    assert (n[0] - n[1]).abs().max().item() <= 2e-2                    # bf16 output rounding only
    assert abs(n.mean().item()) < 2e-2 and n.std().item() == pytest.approx(0.3 / 3 ** 0.5, rel=0.1)
    # Dropout p = 0.5: kept values doubled, about half kept, mask differs across samples
    kept = yd.float() != 0
    assert torch.equal(yd.float()[kept], (2 * xf)[kept])
    assert kept.float().mean().item() == pytest.approx(0.5, abs=0.02)
    assert not torch.equal(kept[0], kept[1])
    # FeatureDropout: pixels whose channel mean reaches u * max are zeroed (threshold ties excluded)
    att = xf.mean(1, keepdim=True)
    thr = att.flatten(1).max(1)[0].view(-1, 1, 1, 1) * u
    margin = (att - thr).abs() > 1e-3 * thr
    expect = xf * (att < thr)
    assert torch.equal((yf.float() * margin), (expect * margin))
    # backward regenerates the same masks
    cots = [torch.randn(shape, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(3)]
    (yn * cots[0]).sum().backward(retain_graph=True)
    gn = x.grad.float().clone(); x.grad = None
    assert torch.allclose(gn, cots[0].float() * (1 + n), rtol=3e-2, atol=3e-2)
    (yd * cots[1]).sum().backward(retain_graph=True)
    gd = x.grad.float().clone(); x.grad = None
    assert torch.allclose(gd, cots[1].float() * kept * 2, rtol=1e-2, atol=1e-2)
    (yf * cots[2]).sum().backward()
    gf = x.grad.float()
    assert torch.equal(gf * margin, cots[2].float() * (att < thr) * margin)


def test_perturb3_nhwc_alias_outputs_fold_every_gradient_into_one_kernel():
    """aliases=2: two extra outputs that ARE x (for the unperturbed consumers); the backward kernel must return the same dx as
    autograd's separate accumulation of the five contributions."""
    from uaps_b200.perturb import perturb3_nhwc
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(3, 32, 24, 16, generator=g, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)      # channels_last [3,16,32,24]
    cots = [torch.randn(3, 32, 24, 16, generator=g, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2) for _ in range(5)]
    xa = x.clone().requires_grad_(True)
    outs = perturb3_nhwc(xa, seed=11, u=0.8, aliases=2)
    assert len(outs) == 5 and outs[3].data_ptr() == xa.data_ptr() and torch.equal(outs[4], xa.detach())
    torch.autograd.backward(outs, cots)
    xb = x.clone().requires_grad_(True)
    ref = perturb3_nhwc(xb, seed=11, u=0.8)
    for o, r in zip(outs[:3], ref):
        assert torch.equal(o, r)
    torch.autograd.backward(ref, cots[:3])
    want = xb.grad.float() + cots[3].float() + cots[4].float()
    assert (xa.grad.float() - want).abs().max().item() <= 2e-2 * want.abs().max().item()     # one bf16 rounding instead of three
    # a single alias and a single perturbed copy (the K = 5 ablation's extra decoder uses outputs=(True, False, False))
    xc = x.clone().requires_grad_(True)
    o1 = perturb3_nhwc(xc, seed=5, u=0.8, outputs=(True, False, False), aliases=1)
    assert o1[1] is None and o1[2] is None and len(o1) == 4
    (o1[0].float().sum() + 2.0 * o1[3].float().sum()).backward()
    assert torch.isfinite(xc.grad.float()).all()
