"""Kernel-only timing (CUDA events around C-ABI calls, no autograd) of the four BatchNorm kernels at the level-0 shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200 import _lib as L
dev = torch.device("cuda:0")
B, H, C = 64, 256, 16
lib = L.lib()
npix = B * H * H
y = torch.randn(npix, C, device=dev).to(torch.bfloat16)
g = torch.randn(npix, C, device=dev).to(torch.bfloat16)
out = torch.empty_like(y); dy = torch.empty_like(y)
sums = torch.zeros(2 * C, dtype=torch.float64, device=dev); sums2 = torch.zeros(2 * C, dtype=torch.float64, device=dev)
stats = torch.empty(2 * C, device=dev); gam = torch.ones(C, device=dev); bet = torch.zeros(C, device=dev)
st = L.stream_ptr()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
mb = npix * C * 2 / 1e6
for p in (0.0, 0.05):
    t_s = timeit(lambda: lib.uaps_bn_stats_nhwc(y.data_ptr(), npix, C, sums.data_ptr(), sums[C:].data_ptr(), st))
    t_a = timeit(lambda: lib.uaps_bn_act_nhwc(y.data_ptr(), sums.data_ptr(), sums[C:].data_ptr(), gam.data_ptr(), bet.data_ptr(), None, None, 0.1, 1e-5, 0.01, p, 7, out.data_ptr(), stats.data_ptr(), stats[C:].data_ptr(), npix, C, None, 1, st))
    t_b = timeit(lambda: lib.uaps_bn_act_bwd_nhwc(g.data_ptr(), y.data_ptr(), gam.data_ptr(), bet.data_ptr(), stats.data_ptr(), stats[C:].data_ptr(), 0.01, p, 7, sums2.data_ptr(), sums2[C:].data_ptr(), dy.data_ptr(), None, None, npix, C, None, st))
    print(f"mult {os.environ.get('UAPS_BN_GRID_MULT','1')} p={p}: stats {t_s:6.1f} us ({mb/t_s*1e3/1e3:5.0f} GB/s) | act {t_a:6.1f} us ({2*mb/t_a:5.2f} TB/s) | bwd pair {t_b:6.1f} us ({5*mb/t_b:5.2f} TB/s)")
x = torch.empty(npix * C, dtype=torch.bfloat16, device=dev)
t_c = timeit(lambda: x.copy_(y.view(-1)))
print(f"torch bf16 copy of the same tensor: {t_c:6.1f} us ({2*mb/t_c:5.2f} TB/s)")
