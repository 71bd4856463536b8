"""torch.profiler view of one bf16 training iteration: which aten ops (torch glue) still launch kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from uaps_b200.train import UAPSTrainer
from uaps_b200.unet import UNet_UAPS
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = UNet_UAPS(3, 4, compute="bf16").to(dev)
tr = UAPSTrainer(model)
xl, xu = torch.randn(B, 3, 256, 256, device=dev), torch.randn(B, 3, 256, 256, device=dev)
yl = torch.randint(0, 4, (B, 256, 256), device=dev)
for _ in range(4):
    tr.step(xl, yl, xu)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False) as prof:
    tr.step(xl, yl, xu)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted(ka, key=lambda e: -e.self_device_time_total)
print(f"{'op':60s} {'calls':>6s} {'self_cuda_us':>12s} {'self_cpu_us':>12s}")
for e in rows[:45]:
    print(f"{e.key[:60]:60s} {e.count:6d} {e.self_device_time_total:12.0f} {e.self_cpu_time_total:12.0f}")
print("--- by CPU time")
for e in sorted(ka, key=lambda e: -e.self_cpu_time_total)[:30]:
    print(f"{e.key[:60]:60s} {e.count:6d} {e.self_device_time_total:12.0f} {e.self_cpu_time_total:12.0f}")
