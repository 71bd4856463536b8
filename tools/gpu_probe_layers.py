"""Per-layer timing of the tcgen05 conv kernels (fprop / dgrad / wgrad) over UNet_UAPS's layer shapes at B images.
Prints time, algorithmic HBM bytes -> GB/s, FLOPs -> TFLOP/s, and the layer's multiplicity in one training iteration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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

tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print(f"{'layer':16s} {'HxW':>5s} {'cin':>7s} {'co':>4s} | {'fprop us':>9s} {'GB/s':>6s} {'TF/s':>6s} | {'dgrad us':>9s} {'GB/s':>6s} | {'wgrad us':>9s} {'GB/s':>6s} {'TF/s':>6s} | x/fwd")
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
    t_f = timeit(lambda: conv(x1, x2))
    t_d = timeit(lambda: dconv(gy, split=c1) if c2 else dconv(gy))
    t_w = timeit(lambda: conv_wgrad(gy, [x1] + ([x2] if c2 else []), co, c1 + c2, ks, out=dw))
    print(f"{name:16s} {H:5d} {c1:3d}+{c2:<3d} {co:4d} | {t_f:9.1f} {by_f / t_f / 1e3:6.0f} {flops / t_f / 1e6:6.1f} | {t_d:9.1f} {by_f / t_d / 1e3:6.0f} | "
          f"{t_w:9.1f} {by_f / t_w / 1e3:6.0f} {flops / t_w / 1e6:6.1f} | {cnt}")
    tot["fprop"] += t_f * cnt * 2; tot["dgrad"] += t_d * cnt * 2; tot["wgrad"] += t_w * cnt * 2
print("per training iteration (2 forwards), ms:", {k: round(v / 1e3, 2) for k, v in tot.items()})
