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


@pytest.mark.parametrize("shape", [(2, 3, 16, 32), (3, 4, 48, 16), (1, 16, 8, 8), (2, 21, 16, 16)])
def test_nchw_f32_to_nhwc_bf16(shape):
    """Entry of the bf16 path: layout + dtype conversion with zero-filled padding channels, bit-exact vs torch."""
    from uaps_b200.conv import pad16, to_nhwc_bf16
    dev = torch.device("cuda:0")
    x = torch.randn(*shape, device=dev)
    out = to_nhwc_bf16(x)
    B, C, H, W = shape
    assert out.shape == (B, H, W, pad16(C)) and out.dtype == torch.bfloat16
    assert torch.equal(out[..., :C], x.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert (out[..., C:] == 0).all()


def test_channel_sums_match_torch():
    from uaps_b200.conv import channel_sums
    dev = torch.device("cuda:0")
    g = torch.randn(4, 32, 24, 64, device=dev).to(torch.bfloat16)          # [B,H,W,C]
    got = channel_sums(g, 50)
    ref = g.double().sum(dim=(0, 1, 2))[:50]
    torch.testing.assert_close(got.double(), ref, rtol=1e-6, atol=1e-4)
