"""One launch of conv_sn_kernel at two shapes (16 -> 16 @256x256, 64 -> 64 @64x64), between cudaProfilerStart/Stop, for
`ncu --set full --import-source on --profile-from-start off`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.conv import PackedConv

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
cases = []
for c, hw in ((16, 256), (64, 64)):
    x = torch.randn(B, hw, hw, c, device=dev).to(torch.bfloat16)
    w = torch.randn(c, c, 3, 3, device=dev) * 0.05
    cases.append((PackedConv(w, torch.zeros(c, device=dev)), x))
for conv, x in cases:
    conv(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for conv, x in cases:
    conv(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
