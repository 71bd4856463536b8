"""GPU parity of the tcgen05 implicit-GEMM convolution against torch's conv2d on the same bf16-rounded
operands (fp32 accumulation both sides).  Tolerance 1e-2 of the output scale, as BASELINE.json's
north_star states for bf16 conv activations; observed errors are ~1e-3 (bf16 output rounding)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, w, b, ks):
    xr, wr = x.bfloat16().float(), w.bfloat16().float()
    return F.conv2d(xr, wr, b, padding=ks // 2)


# (B, H, W, Cin, Cout, ks): every distinct layer shape of UNet_UAPS at 256x256 plus edge cases
SHAPES = [
    (2, 64, 64, 3, 16, 3),       # in_conv first conv (Cin padded 3 -> 16)
    (2, 64, 64, 16, 16, 3),      # level-0 blocks
    (2, 32, 32, 16, 32, 3),      # down1
    (2, 32, 32, 32, 32, 3),
    (2, 16, 16, 32, 64, 3),      # down2
    (2, 16, 16, 64, 64, 3),
    (1, 32, 32, 64, 128, 3),     # down3
    (1, 32, 32, 128, 128, 3),
    (1, 16, 16, 128, 256, 3),    # down4: two N tiles
    (1, 16, 16, 256, 256, 3),
    (1, 16, 16, 256, 128, 1),    # up1.conv1x1
    (2, 32, 32, 32, 16, 1),      # up4.conv1x1
    (2, 64, 64, 16, 4, 3),       # out_conv (N padded 4 -> 16)
    (1, 48, 40, 16, 16, 3),      # partial tiles in y (48 = 3 x 16) and x (40 = 5 x 8)
    (1, 15, 40, 128, 128, 3),    # KoSDD2 240x640 at level 4: H not a multiple of the tile
    (1, 24, 20, 32, 32, 3),      # W not a multiple of 8
]


@pytest.mark.parametrize("B,H,W,ci,co,ks", SHAPES)
def test_fprop_matches_torch(B, H, W, ci, co, ks):
    from uaps_b200.conv import PackedConv, from_nhwc, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 1000 + H + ci * 7 + co)
    x = torch.randn(B, ci, H, W, generator=g).to(dev)
    w = (torch.randn(co, ci, ks, ks, generator=g) * (2.0 / (ci * ks * ks)) ** 0.5).to(dev)
    b = torch.randn(co, generator=g).to(dev)
    ref = _ref(x, w, b, ks)
    conv = PackedConv(w, b)
    y = conv(to_nhwc_bf16(x))
    out = from_nhwc(y, co)
    err = (out - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item(), (err, ref.abs().max().item())
    if y.shape[-1] > co:
        assert y[..., co:].abs().max().item() == 0.0          # padding channels stay zero


def test_concat_two_segments():
    """conv(cat([skip, up])) as two K segments (UpBlock, UAPS_unet.py:85-86) without the concat."""
    from uaps_b200.conv import PackedConv, from_nhwc, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    for c in (16, 32, 64, 128):
        skip, up = torch.randn(2, c, 32, 32, generator=g).to(dev), torch.randn(2, c, 32, 32, generator=g).to(dev)
        w = (torch.randn(c, 2 * c, 3, 3, generator=g) * (1.0 / (18 * c)) ** 0.5).to(dev)
        b = torch.randn(c, generator=g).to(dev)
        ref = _ref(torch.cat([skip, up], 1), w, b, 3)
        out = from_nhwc(PackedConv(w, b, cin_split=c)(to_nhwc_bf16(skip), to_nhwc_bf16(up)), c)
        assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item(), c


def test_logits_output_nchw_fp32():
    """out_conv writes fp32 NCHW directly -- the layout the fused loss kernel reads."""
    from uaps_b200.conv import PackedConv, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(6)
    for C in (2, 4, 7):
        x = torch.randn(2, 16, 64, 64, generator=g).to(dev)
        w, b = (torch.randn(C, 16, 3, 3, generator=g) * 0.1).to(dev), torch.randn(C, generator=g).to(dev)
        ref = _ref(x, w, b, 3)
        out = PackedConv(w, b)(to_nhwc_bf16(x), out_nchw_f32=True)
        assert out.shape == ref.shape and out.dtype == torch.float32
        assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()    # fp32 output: only operand rounding


def test_data_gradient_via_transposed_packing():
    """dX of a 3x3 conv = conv of dY with the rotated, channel-transposed kernel (same tcgen05 kernel)."""
    from uaps_b200.conv import PackedConv, from_nhwc, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(8)
    for ci, co in ((16, 32), (64, 64), (128, 64)):
        x = torch.randn(2, ci, 32, 32, generator=g).to(dev).requires_grad_(True)
        w = (torch.randn(co, ci, 3, 3, generator=g) * 0.05).to(dev)
        dy = torch.randn(2, co, 32, 32, generator=g).to(dev)
        F.conv2d(x, w.bfloat16().float(), padding=1).backward(dy.bfloat16().float())
        dx = from_nhwc(PackedConv(w, None, transpose=True)(to_nhwc_bf16(dy)), ci)
        assert (dx - x.grad).abs().max().item() <= 1e-2 * x.grad.abs().max().item(), (ci, co)


@pytest.mark.parametrize("B,H,W,ci,co,ks", [
    (2, 32, 32, 16, 16, 3), (2, 32, 32, 32, 32, 3), (2, 16, 16, 64, 64, 3), (1, 32, 32, 128, 128, 3),
    (1, 16, 16, 256, 256, 3), (2, 32, 32, 32, 16, 1), (1, 16, 16, 256, 128, 1), (2, 64, 64, 16, 4, 3),
    (1, 48, 40, 32, 64, 3), (1, 15, 40, 128, 128, 3), (4, 64, 64, 16, 32, 3),
])
def test_weight_gradient_matches_torch(B, H, W, ci, co, ks):
    """tcgen05 wgrad (MN-major operands) vs autograd of torch's conv on the same bf16-rounded tensors."""
    from uaps_b200.conv import conv_wgrad, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B + H + ci + 3 * co)
    x = torch.randn(B, ci, H, W, generator=g).to(dev)
    dy = torch.randn(B, co, H, W, generator=g).to(dev)
    w = torch.zeros(co, ci, ks, ks, device=dev, requires_grad=True)
    F.conv2d(x.bfloat16().float(), w, padding=ks // 2).backward(dy.bfloat16().float())
    dw = conv_wgrad(to_nhwc_bf16(dy), [to_nhwc_bf16(x)], co, ci, ks)
    err = (dw - w.grad).abs().max().item()
    assert err <= 2e-3 * w.grad.abs().max().item(), (err, w.grad.abs().max().item())


def test_weight_gradient_of_concat_input():
    from uaps_b200.conv import conv_wgrad, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(12)
    for c in (16, 64, 128):
        skip, up = torch.randn(2, c, 32, 32, generator=g).to(dev), torch.randn(2, c, 32, 32, generator=g).to(dev)
        dy = torch.randn(2, c, 32, 32, generator=g).to(dev)
        w = torch.zeros(c, 2 * c, 3, 3, device=dev, requires_grad=True)
        F.conv2d(torch.cat([skip, up], 1).bfloat16().float(), w, padding=1).backward(dy.bfloat16().float())
        dw = conv_wgrad(to_nhwc_bf16(dy), [to_nhwc_bf16(skip), to_nhwc_bf16(up)], c, 2 * c, 3)
        assert (dw - w.grad).abs().max().item() <= 2e-3 * w.grad.abs().max().item(), c


def test_split_outputs_of_concat_data_gradient():
    """The data gradient of a concat conv is written straight into two tensors (channels [0,c) and [c,2c))."""
    from uaps_b200.conv import PackedConv, from_nhwc, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(21)
    for c in (16, 64):
        w = (torch.randn(c, 2 * c, 3, 3, generator=g) * 0.05).to(dev)
        dy = torch.randn(2, c, 32, 32, generator=g).to(dev)
        conv = PackedConv(w, None, transpose=True)
        whole = conv(to_nhwc_bf16(dy))
        a, b = conv(to_nhwc_bf16(dy), split=c)
        assert a.shape[-1] == c and b.shape[-1] == c
        assert torch.equal(a, whole[..., :c]) and torch.equal(b, whole[..., c:])


@pytest.mark.parametrize("B,H,W,ci,co,ks,fold", [
    (2, 64, 64, 16, 16, 3, 4), (2, 64, 64, 3, 16, 3, 4), (2, 32, 64, 16, 32, 3, 2), (2, 32, 32, 32, 32, 3, 2),
    (1, 32, 64, 32, 64, 3, 2), (2, 64, 64, 32, 16, 1, 2), (2, 48, 96, 16, 16, 3, 4), (1, 16, 32, 16, 16, 3, 4),
])
def test_pixel_folded_conv_matches_torch(B, H, W, ci, co, ks, fold):
    """Pixel folding (F pixels viewed as one pixel with F*C channels, block-banded weights) gives the same conv."""
    from uaps_b200.conv import PackedConv, from_nhwc, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B + H + ci + co + fold)
    x = torch.randn(B, ci, H, W, generator=g).to(dev)
    w = (torch.randn(co, ci, ks, ks, generator=g) * (2.0 / (ci * ks * ks)) ** 0.5).to(dev)
    b = torch.randn(co, generator=g).to(dev)
    ref = _ref(x, w, b, ks)
    out = from_nhwc(PackedConv(w, b, fold=fold)(to_nhwc_bf16(x)), co)
    assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    # and identical (up to accumulation order) to the unfolded kernel
    plain = from_nhwc(PackedConv(w, b, fold=1)(to_nhwc_bf16(x)), co)
    assert (out - plain).abs().max().item() <= 4e-3 * ref.abs().max().item()


def test_pixel_folded_variants():
    """Folded concat (two K segments), folded data gradient with split outputs, folded fp32-NCHW logits."""
    from uaps_b200.conv import PackedConv, from_nhwc, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(31)
    c = 16
    skip, up = torch.randn(2, c, 32, 64, generator=g).to(dev), torch.randn(2, c, 32, 64, generator=g).to(dev)
    w = (torch.randn(c, 2 * c, 3, 3, generator=g) * 0.06).to(dev)
    b = torch.randn(c, generator=g).to(dev)
    ref = _ref(torch.cat([skip, up], 1), w, b, 3)
    out = from_nhwc(PackedConv(w, b, cin_split=c, fold=2)(to_nhwc_bf16(skip), to_nhwc_bf16(up)), c)
    assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    dy = torch.randn(2, c, 32, 64, generator=g).to(dev)
    whole = PackedConv(w, None, transpose=True, fold=1)(to_nhwc_bf16(dy))
    a, bb = PackedConv(w, None, transpose=True, fold=2)(to_nhwc_bf16(dy), split=c)
    assert (a.float() - whole[..., :c].float()).abs().max().item() <= 2e-2 * whole.float().abs().max().item()
    assert (bb.float() - whole[..., c:].float()).abs().max().item() <= 2e-2 * whole.float().abs().max().item()
    x = torch.randn(2, 16, 32, 64, generator=g).to(dev)
    wl, bl = (torch.randn(4, 16, 3, 3, generator=g) * 0.1).to(dev), torch.randn(4, generator=g).to(dev)
    refl = _ref(x, wl, bl, 3)
    outl = PackedConv(wl, bl, fold=4)(to_nhwc_bf16(x), out_nchw_f32=True)
    assert outl.shape == refl.shape and (outl - refl).abs().max().item() <= 2e-3 * refl.abs().max().item()


@pytest.mark.parametrize("shape", [(4, 64, 64, 16, 16, 3), (2, 32, 32, 128, 128, 3), (3, 16, 16, 256, 256, 3), (2, 64, 64, 64, 32, 1),
                                   (2, 48, 80, 32, 64, 3), (8, 128, 128, 16, 4, 3)])
def test_wgrad_deterministic_split_k(shape):
    """Deterministic split-K (workspace + in-launch fold, no atomics): bit-reproducible, equal to the atomic path up to
    fp32 re-association, accumulates into an existing gradient, and reuses one workspace across layers."""
    from uaps_b200.conv import conv_wgrad
    B, H, W, cin, cout, ks = shape
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(B * H + cin)
    x = torch.randn(B, H, W, cin, generator=g, device=dev).to(torch.bfloat16)
    cop = (cout + 15) // 16 * 16
    dy = torch.randn(B, H, W, cop, generator=g, device=dev).to(torch.bfloat16)
    dy[..., cout:] = 0
    a = conv_wgrad(dy, [x], cout, cin, ks, deterministic=True)
    b = conv_wgrad(dy, [x], cout, cin, ks, deterministic=True)
    if cout * cin * ks * ks >= 16384:                     # smaller layers keep the (cheap) atomic epilogue, see conv_wgrad.cu
        assert torch.equal(a, b), "deterministic split-K must be bit-reproducible"
    c = conv_wgrad(dy, [x], cout, cin, ks, deterministic=False)
    torch.testing.assert_close(a, c, rtol=1e-4, atol=1e-4 * c.abs().max().item())
    # fp32 reference on the same bf16 operands
    xr, dyr = x.float().permute(0, 3, 1, 2), dy[..., :cout].float().permute(0, 3, 1, 2)
    w = torch.zeros(cout, cin, ks, ks, device=dev, requires_grad=True)
    (torch.nn.functional.conv2d(xr, w, padding=ks // 2) * dyr).sum().backward()
    assert (a - w.grad).abs().max().item() <= 2e-3 * w.grad.abs().max().item() + 1e-5
    base = torch.full_like(a, 0.5)                       # accumulates into what is there
    conv_wgrad(dy, [x], cout, cin, ks, out=base, deterministic=True)
    torch.testing.assert_close(base, a + 0.5, rtol=1e-4, atol=1e-4 * a.abs().max().item() + 1e-6)   # (atomic path: re-association)


@pytest.mark.parametrize("shape", [(4, 64, 64, 16, 0, 16, 3), (2, 32, 48, 32, 32, 32, 3), (2, 16, 16, 128, 0, 256, 3), (3, 40, 24, 64, 0, 64, 3),
                                   (2, 16, 16, 256, 0, 128, 1)])
def test_conv_epilogue_batchnorm_statistics(shape):
    """uaps_conv_fprop_bn: the per-channel sum / sum of squares of the conv output, accumulated by the epilogue into
    replicated fp64 sums, against the statistics of the bf16 output tensor the same call writes."""
    from uaps_b200.conv import PackedConv
    B, H, W, c1, c2, co, ks = shape
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(H * W + co)
    x1 = torch.randn(B, H, W, c1, generator=g, device=dev).to(torch.bfloat16)
    x2 = torch.randn(B, H, W, c2, generator=g, device=dev).to(torch.bfloat16) if c2 else None
    w = torch.randn(co, c1 + c2, ks, ks, generator=g, device=dev) * 0.1
    b = torch.randn(co, generator=g, device=dev)
    conv = PackedConv(w, b, cin_split=c1 if c2 else None)
    nrep = 8
    sums = torch.zeros(nrep * 2 * co, dtype=torch.float64, device=dev)
    y = conv(x1, x2, bn_sums=sums, bn_nrep=nrep)
    y_plain = conv(x1, x2)
    assert torch.equal(y, y_plain)                                   # the statistics do not disturb the output
    tot = sums.view(nrep, 2, co).sum(0)
    yf = y.double().view(-1, co)
    n = yf.shape[0]
    mean_ref, var_ref = yf.mean(0), yf.var(0, unbiased=False)
    mean, var = tot[0] / n, tot[1] / n - (tot[0] / n) ** 2
    # the epilogue sums the fp32 accumulators, the reference the bf16-rounded tensor: unbiased rounding noise of 2^-9 per element
    assert (mean - mean_ref).abs().max().item() <= 2e-3 * yf.abs().mean().item()
    assert ((var - var_ref).abs() / var_ref).max().item() <= 2e-3
    assert (sums.view(nrep, 2, co)[:, 0].abs().sum(1) > 0).sum().item() >= min(nrep, 2)      # the replicas are really used


# conv_sn_kernel (3x3 layers with <= 64 output channels and input channels >= 2 x output channels or >= 64: the horizontal
# taps sit in the MMA's N dimension, the pixel shift is done by the epilogue, and the shifts that cross a 32-pixel tile edge
# are carried from tile to tile along a 4-row band): widths with several tiles per band, a ragged last tile, a last tile that
# ends exactly on lane 31, ragged bands, every epilogue variant, every channel-chunk width and epilogue-warp count.
SN_SHAPES = [
    (2, 10, 100, 32, 16),      # 4 tiles per band (100 = 3 x 32 + 4), H = 10: a 2-row last band
    (1, 7, 96, 64, 32),        # exactly 3 tiles: the band's last pixel is lane 31 of the last tile
    (2, 9, 33, 96, 48),        # one pixel in the second tile; n_co = 48 -> N = 144, three chunks of 32
    (1, 12, 160, 64, 64),
    (1, 5, 72, 128, 64),       # two channel chunks of 64 (data gradient of down3's first conv)
    (2, 8, 64, 48, 16),        # three chunks of 16
    (1, 6, 40, 64, 4),         # 4 real output channels in a 16-column group
    (1, 4, 31, 80, 48),        # five chunks of 16; a single ragged tile
    # conv_sn_small_kernel (<= 5 output channels: the logits layer; all taps of all classes in one 16-column group)
    (2, 9, 100, 16, 2),
    (1, 7, 96, 32, 3),
    (2, 5, 33, 16, 5),
    (1, 8, 64, 16, 1),
    (2, 6, 72, 16, 4),
]


@pytest.mark.parametrize("B,H,W,ci,co", SN_SHAPES)
def test_sn_kernel_tile_edges(B, H, W, ci, co):
    from uaps_b200.conv import PackedConv, from_nhwc, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(H * 131 + W + ci + 3 * co)
    x = torch.randn(B, ci, H, W, generator=g).to(dev)
    w = (torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (ci * 9)) ** 0.5).to(dev)
    b = torch.randn(co, generator=g).to(dev)
    ref = _ref(x, w, b, 3)
    scale = ref.abs().max().item()
    conv = PackedConv(w, b)
    xn = to_nhwc_bf16(x)
    out = from_nhwc(conv(xn), co)
    err = (out - ref).abs()
    assert err.max().item() <= 1e-2 * scale, (err.max().item(), scale, torch.nonzero(err > 1e-2 * scale)[:8].tolist())
    # LeakyReLU epilogue (inference with BatchNorm folded in)
    out = from_nhwc(conv(xn, slope=0.01), co)
    assert (out - F.leaky_relu(ref, 0.01)).abs().max().item() <= 1e-2 * scale
    # fp32 NCHW epilogue (logits)
    out = conv(xn, out_nchw_f32=True)
    assert (out - ref).abs().max().item() <= 2e-3 * scale
    # BatchNorm statistics epilogue: same output, sums of exactly the pixels of the image
    sums = torch.zeros(4 * 2 * ((co + 15) // 16 * 16), dtype=torch.float64, device=dev)
    y = conv(xn, bn_sums=sums, bn_nrep=4)
    assert torch.equal(y, conv(xn))
    cp = (co + 15) // 16 * 16
    tot = sums.view(4, 2, cp).sum(0)[:, :co]
    yf = y[..., :co].double().reshape(-1, co)
    assert (tot[0] - yf.sum(0)).abs().max().item() <= 3e-3 * yf.abs().sum(0).max().item()
    assert ((tot[1] - (yf * yf).sum(0)).abs() / (yf * yf).sum(0)).max().item() <= 3e-3


def test_sn_kernel_concat_and_split_over_several_tiles():
    from uaps_b200.conv import PackedConv, from_nhwc, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(77)
    for c, H, W in ((16, 9, 100), (32, 6, 70), (64, 5, 64)):
        skip, up = torch.randn(2, c, H, W, generator=g).to(dev), torch.randn(2, c, H, W, generator=g).to(dev)
        w = (torch.randn(c, 2 * c, 3, 3, generator=g) * (1.0 / (18 * c)) ** 0.5).to(dev)
        b = torch.randn(c, generator=g).to(dev)
        ref = _ref(torch.cat([skip, up], 1), w, b, 3)
        out = from_nhwc(PackedConv(w, b, cin_split=c)(to_nhwc_bf16(skip), to_nhwc_bf16(up)), c)
        assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item(), c
    for c, co, H, W in ((16, 64, 9, 100), (32, 128, 6, 70)):  # data gradient of a concat conv, written into two tensors
        w = (torch.randn(co, 2 * c, 3, 3, generator=g) * 0.05).to(dev)
        dy = torch.randn(2, co, H, W, generator=g).to(dev)
        x = torch.zeros(2, 2 * c, H, W, device=dev, requires_grad=True)
        F.conv2d(x, w.bfloat16().float(), padding=1).backward(dy.bfloat16().float())
        conv = PackedConv(w, None, transpose=True)
        whole = conv(to_nhwc_bf16(dy))
        a, b2 = conv(to_nhwc_bf16(dy), split=c)
        assert torch.equal(a, whole[..., :c]) and torch.equal(b2, whole[..., c:])
        assert (from_nhwc(whole, 2 * c) - x.grad).abs().max().item() <= 1e-2 * x.grad.abs().max().item()


@pytest.mark.parametrize("B,C,H,W", [(3, 4, 40, 56), (2, 2, 64, 64), (1, 7, 17, 33), (2, 8, 16, 16)])
def test_entry_conversion_with_bias_gradient_sums(B, C, H, W):
    """uaps_nchw_f32_to_nhwc_bf16_sums: the same tensor as the plain conversion, and the per-channel sums of exactly the
    bf16 values it holds (out_conv's bias gradient without a separate reduction pass)."""
    from uaps_b200.conv import to_nhwc_bf16, to_nhwc_bf16_with_sums
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(B * 100 + C)
    x = torch.randn(B, C, H, W, generator=g, device=dev) * 3.0
    ref = to_nhwc_bf16(x)
    out, sums = to_nhwc_bf16_with_sums(x)
    assert torch.equal(out, ref)
    want = ref[..., :C].double().sum((0, 1, 2))
    assert sums.shape == (C,) and sums.dtype == torch.float32
    torch.testing.assert_close(sums.double(), want, rtol=1e-5, atol=1e-4)


def test_logits_conv_bias_gradient():
    """out_conv through autograd: weight, bias and input gradients against fp32 torch on the bf16-rounded operands."""
    from uaps_b200.conv import conv_bf16, to_nhwc_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 16, 48, 40, generator=g).to(dev)
    w = (torch.randn(4, 16, 3, 3, generator=g) * 0.1).to(dev).requires_grad_(True)
    b = torch.randn(4, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn(2, 4, 48, 40, generator=g).to(dev)
    xc = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = conv_bf16(xc, w, b, nchw_f32_out=True)
    y.backward(gy)
    xr = x.bfloat16().float().requires_grad_(True)
    wr, br = w.detach().bfloat16().float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    F.conv2d(xr, wr, br, padding=1).backward(gy)
    assert (b.grad - br.grad).abs().max().item() <= 2e-2 * br.grad.abs().max().item()
    assert (w.grad - wr.grad).abs().max().item() <= 1e-2 * wr.grad.abs().max().item()
    assert (xc.grad.float() - xr.grad).abs().max().item() <= 1e-2 * xr.grad.abs().max().item()
