"""Device timing of the fused BatchNorm + LeakyReLU + dropout kernels (forward pair, backward pair) at UNet_UAPS shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.bn_act import bn_lrelu_dropout
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print(f"variant {os.environ.get('UAPS_BN_BWD_VARIANT', 'default')}")
for (H, C, p) in [(256, 16, 0.0), (256, 16, 0.05), (128, 32, 0.0), (64, 64, 0.0), (32, 128, 0.0)]:
    bn = torch.nn.BatchNorm2d(C).to(dev)
    y = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
    g = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
    t_f = timeit(lambda: bn_lrelu_dropout(y, bn, p, seed=7))
    a = bn_lrelu_dropout(y, bn, p, seed=7)
    t_b = timeit(lambda: torch.autograd.grad(a, [y], g, retain_graph=True))
    mb = B * H * H * C * 2 / 1e6
    print(f"H {H:4d} C {C:4d} p {p:.2f}: fwd (stats + act) {t_f:7.1f} us = {3 * mb / t_f * 1e3 / 1e3:6.0f} GB/s | bwd (reduce + apply) {t_b:7.1f} us = {5 * mb / t_b * 1e3 / 1e3:6.0f} GB/s")
