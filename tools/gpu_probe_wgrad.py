import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.conv import conv_wgrad, to_nhwc_bf16
dev = "cuda:0"
def timeit(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (B, H, W, ci, co) in [(16, 256, 256, 16, 16), (16, 32, 32, 128, 128), (16, 64, 64, 64, 64), (16, 128, 128, 32, 32), (16, 16, 16, 256, 256)]:
    x = to_nhwc_bf16(torch.randn(B, ci, H, W, device=dev)); dy = to_nhwc_bf16(torch.randn(B, co, H, W, device=dev))
    t = timeit(lambda: conv_wgrad(dy, [x], co, ci, 3))
    fl = 2.0 * B * H * W * co * ci * 9
    print(f"wgrad B={B} {H}x{W} {ci}->{co}: {t*1e3:.1f} us  {fl/t/1e9:.1f} TF/s", flush=True)
