"""One bf16 training iteration at 16+16 images for kernel-level profiling (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.train import UAPSTrainer
from uaps_b200.unet import UNet_UAPS
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = UNet_UAPS(3, 4, compute="bf16").to(dev)
tr = UAPSTrainer(model)
xl, xu = torch.randn(B, 3, 256, 256, device=dev), torch.randn(B, 3, 256, 256, device=dev)
yl = torch.randint(0, 4, (B, 256, 256), device=dev)
for _ in range(4):
    out = tr.step(xl, yl, xu)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    out = tr.step(xl, yl, xu)
torch.cuda.synchronize()
print("wall ms/iter", (time.perf_counter() - t0) / 5 * 1e3)
torch.cuda.profiler.start()
out = tr.step(xl, yl, xu)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", out["loss"].item())
