"""Per-layer timing of the tcgen05 conv kernels (fprop / dgrad / wgrad) over UNet_UAPS's layer shapes at B images.
Prints time, algorithmic HBM bytes -> GB/s, FLOPs -> TFLOP/s, and the layer's multiplicity in one training iteration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from uaps_b200.conv import PackedConv, conv_wgrad, pad16
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
# (name, H, cin1, cin2, cout, ks, count per forward)
LAYERS = [("enc0.c1", 256, 16, 0, 16, 3, 1), ("enc0.c2/up4.c2", 256, 16, 0, 16, 3, 5), ("up4.c1 cat", 256, 16, 16, 16, 3, 4),
          ("out_conv", 256, 16, 0, 4, 3, 4), ("enc1.c1", 128, 16, 0, 32, 3, 1), ("enc1.c2/up3.c2", 128, 32, 0, 32, 3, 5),
          ("up3.c1 cat", 128, 32, 32, 32, 3, 4), ("up4.1x1", 128, 32, 0, 16, 1, 4), ("enc2.c1", 64, 32, 0, 64, 3, 1),
          ("enc2.c2/up2.c2", 64, 64, 0, 64, 3, 5), ("up2.c1 cat", 64, 64, 64, 64, 3, 4), ("up3.1x1", 64, 64, 0, 32, 1, 4),
          ("enc3.c1", 32, 64, 0, 128, 3, 1), ("enc3.c2/up1.c2", 32, 128, 0, 128, 3, 5), ("up1.c1 cat", 32, 128, 128, 128, 3, 4),
          ("up2.1x1", 32, 128, 0, 64, 1, 4), ("enc4.c1", 16, 128, 0, 256, 3, 1), ("enc4.c2", 16, 256, 0, 256, 3, 1),
          ("up1.1x1", 16, 256, 0, 128, 1, 4)]

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0, "cudnn_fprop": 0.0, "cudnn_dgrad": 0.0, "cudnn_wgrad": 0.0}
print("uaps = hand-written tcgen05 kernels; cuDNN = torch bf16 channels_last on the same box (F.conv2d / aten.convolution_backward;")
print("for the concat layers cuDNN is given the ALREADY concatenated tensor, i.e. the torch.cat pass is not charged to it)")
print(f"{'layer':16s} {'HxW':>5s} {'cin':>7s} {'co':>4s} | {'fprop us':>9s} {'GB/s':>6s} {'TF/s':>6s} {'cuDNN':>7s} | {'dgrad us':>9s} {'GB/s':>6s} {'cuDNN':>7s} | {'wgrad us':>9s} {'GB/s':>6s} {'TF/s':>6s} {'cuDNN':>7s} | x/fwd")
for name, H, c1, c2, co, ks, cnt in LAYERS:
    W = H
    x1 = torch.randn(B, H, W, c1, device=dev).to(torch.bfloat16)
    x2 = torch.randn(B, H, W, c2, device=dev).to(torch.bfloat16) if c2 else None
    w = torch.randn(co, c1 + c2, ks, ks, device=dev) * 0.05
    b = torch.zeros(co, device=dev)
    conv = PackedConv(w, b, cin_split=c1 if c2 else None)
    dconv = PackedConv(w, None, transpose=True)
    cop = pad16(co)
    gy = torch.randn(B, H, W, cop, device=dev).to(torch.bfloat16)
    if cop != co: gy[..., co:] = 0
    dw = torch.zeros(co, c1 + c2, ks, ks, device=dev)
    npix = B * H * W
    flops = 2.0 * npix * co * (c1 + c2) * ks * ks
    by_f = npix * 2 * (c1 + c2 + cop)
    out_kw = {"out_nchw_f32": True} if name == "out_conv" else {}        # the logits layer writes fp32 NCHW, as in the model
    t_f = timeit(lambda: conv(x1, x2, **out_kw))
    # the library's bf16 kernels on the same operands
    xt = (torch.cat([x1, x2], 3) if c2 else x1).permute(0, 3, 1, 2)
    w16 = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b16 = b.to(torch.bfloat16)
    gyc = gy[..., :co].permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    pd = ks // 2
    c_f = timeit(lambda: F.conv2d(xt, w16, b16, padding=pd))
    c_d = timeit(lambda: torch.ops.aten.convolution_backward(gyc, xt, w16, None, [1, 1], [pd, pd], [1, 1], False, [0, 0], 1, [True, False, False]))
    c_w = timeit(lambda: torch.ops.aten.convolution_backward(gyc, xt, w16, None, [1, 1], [pd, pd], [1, 1], False, [0, 0], 1, [False, True, False]))
    t_d = timeit(lambda: dconv(gy, split=c1) if c2 else dconv(gy))
    t_w = timeit(lambda: conv_wgrad(gy, [x1] + ([x2] if c2 else []), co, c1 + c2, ks, out=dw))
    print(f"{name:16s} {H:5d} {c1:3d}+{c2:<3d} {co:4d} | {t_f:9.1f} {by_f / t_f / 1e3:6.0f} {flops / t_f / 1e6:6.1f} {c_f:7.1f} | {t_d:9.1f} {by_f / t_d / 1e3:6.0f} {c_d:7.1f} | "
          f"{t_w:9.1f} {by_f / t_w / 1e3:6.0f} {flops / t_w / 1e6:6.1f} {c_w:7.1f} | {cnt}")
    tot["fprop"] += t_f * cnt * 2; tot["dgrad"] += t_d * cnt * 2; tot["wgrad"] += t_w * cnt * 2
    tot["cudnn_fprop"] += c_f * cnt * 2; tot["cudnn_dgrad"] += c_d * cnt * 2; tot["cudnn_wgrad"] += c_w * cnt * 2
print("per training iteration (2 forwards), ms:", {k: round(v / 1e3, 2) for k, v in tot.items()})
