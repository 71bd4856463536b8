import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from uaps_b200.unet import UNet_UAPS
from uaps_b200.train import UAPSTrainer, UAPSConfig
import uaps_b200.unet as U
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 8
for graph in (False, True):
    model = UNet_UAPS(3, 4, compute="bf16").to(dev)
    tr = UAPSTrainer(model, UAPSConfig(cuda_graph=graph))
    xl, xu = torch.randn(B, 3, 256, 256, device=dev), torch.randn(B, 3, 256, 256, device=dev)
    yl = torch.randint(0, 4, (B, 256, 256), device=dev)
    for _ in range(4): tr.step(xl, yl, xu)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.step(xl, yl, xu); torch.cuda.synchronize()
    print("graph", graph, "model.training", model.training)
    for e in prof.key_averages():
        if any(k in e.key for k in ("bn_stats", "bn_act_kernel", "conv_igemm", "conv_sn", "wgrad")):
            print(f"   {e.key[:80]:80s} {e.count:5d} {e.self_device_time_total:9.0f}")
