"""GPU parity of the channels-last bf16 upsample / max-pool kernels against torch (same bf16 input)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("B,C,H,W", [(2, 128, 16, 16), (2, 16, 64, 64), (1, 32, 15, 40), (2, 64, 2, 2), (1, 8, 1, 3)])
def test_upsample2x(B, C, H, W):
    from uaps_b200.resample import upsample2x
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(H * W + C)
    x = _cl(torch.randn(B, C, H, W, generator=g).to(dev))
    cot = torch.randn(B, C, 2 * H, 2 * W, generator=g).to(dev)
    xr = x.float().requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    (ref * cot.bfloat16().float()).sum().backward()
    xo = x.clone().requires_grad_(True)
    out = upsample2x(xo)
    (out.float() * cot.bfloat16().float()).sum().backward()
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert (xo.grad.float() - xr.grad).abs().max().item() <= 1e-2 * xr.grad.abs().max().item()


@pytest.mark.parametrize("B,C,H,W", [(2, 16, 64, 64), (2, 128, 8, 8), (1, 32, 30, 80)])
def test_maxpool2(B, C, H, W):
    from uaps_b200.resample import maxpool2
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(H + C)
    x = _cl(torch.randn(B, C, H, W, generator=g).to(dev))
    cot = _cl(torch.randn(B, C, H // 2, W // 2, generator=g).to(dev))
    xr = x.float().requires_grad_(True)
    ref = F.max_pool2d(xr, 2)
    (ref * cot.float()).sum().backward()
    xo = x.clone().requires_grad_(True)
    out = maxpool2(xo)
    (out.float() * cot.float()).sum().backward()
    assert torch.equal(out.float(), ref)
    assert torch.equal(xo.grad.float(), xr.grad)
