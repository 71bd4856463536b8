"""Device timing of the tcgen05 conv against cuDNN (bf16 channels_last) per UNet_UAPS layer shape (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from uaps_b200.conv import PackedConv, pick_fold, to_nhwc_bf16

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def run(B, H, W, ci, co, ks, split=None, name=""):
    dev = "cuda:0"
    x = torch.randn(B, ci, H, W, device=dev)
    w = torch.randn(co, ci, ks, ks, device=dev) * 0.05
    b = torch.randn(co, device=dev)
    fold = pick_fold([ci] if split is None else [split, ci - split], co, ks, W)
    conv = PackedConv(w, b, cin_split=split, fold=fold)
    if split is None:
        xs = (to_nhwc_bf16(x),)
    else:
        xs = (to_nhwc_bf16(x[:, :split]), to_nhwc_bf16(x[:, split:]))
    t_ours = timeit(lambda: conv(*xs))
    xc = x.bfloat16().contiguous(memory_format=torch.channels_last)
    wc = w.bfloat16().contiguous(memory_format=torch.channels_last)
    bc = b.bfloat16()
    t_cudnn = timeit(lambda: F.conv2d(xc, wc, bc, padding=ks // 2))
    flop = 2.0 * B * H * W * co * ci * ks * ks
    byts = B * H * W * (max(ci, 16) + max(co, 16)) * 2
    print(f"{name:14s} F={fold} B={B} {H}x{W} {ci}->{co} k{ks}: ours {t_ours*1e3:8.1f}us {flop/t_ours/1e9:8.1f} TF/s {byts/t_ours/1e6:7.0f} GB/s | "
          f"cuDNN {t_cudnn*1e3:8.1f}us {flop/t_cudnn/1e9:8.1f} TF/s | speedup {t_cudnn/t_ours:5.2f}x", flush=True)

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    run(B, 256, 256, 3, 16, 3, name="enc0.conv1")
    run(B, 256, 256, 16, 16, 3, name="enc0.conv2")
    run(B, 128, 128, 16, 32, 3, name="enc1.conv1")
    run(B, 128, 128, 32, 32, 3, name="enc1.conv2")
    run(B, 64, 64, 32, 64, 3, name="enc2.conv1")
    run(B, 64, 64, 64, 64, 3, name="enc2.conv2")
    run(B, 32, 32, 64, 128, 3, name="enc3.conv1")
    run(B, 32, 32, 128, 128, 3, name="enc3.conv2")
    run(B, 16, 16, 128, 256, 3, name="enc4.conv1")
    run(B, 16, 16, 256, 256, 3, name="enc4.conv2")
    run(B, 16, 16, 256, 128, 1, name="up1.conv1x1")
    run(B, 32, 32, 256, 128, 3, split=128, name="up1.conv1")
    run(B, 32, 32, 128, 128, 3, name="up1.conv2")
    run(B, 64, 64, 128, 64, 3, split=64, name="up2.conv1")
    run(B, 128, 128, 64, 32, 3, split=32, name="up3.conv1")
    run(B, 256, 256, 32, 16, 3, split=16, name="up4.conv1")
    run(B, 256, 256, 16, 4, 3, name="out_conv")
