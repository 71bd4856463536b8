"""One launch of conv_sn_kernel at the shapes where make_plan selects it (concat 16+16 -> 16 @256x256, concat 32+32 -> 32
@128x128, 64 -> 64 @64x64), between cudaProfilerStart/Stop, for `ncu --set full --import-source on --profile-from-start off`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.conv import PackedConv

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
cases = []
for c, hw, cat in ((16, 256, True), (32, 128, True), (64, 64, False)):
    x = torch.randn(B, hw, hw, c, device=dev).to(torch.bfloat16)
    x2 = torch.randn(B, hw, hw, c, device=dev).to(torch.bfloat16) if cat else None
    w = torch.randn(c, 2 * c if cat else c, 3, 3, device=dev) * 0.05
    cases.append((PackedConv(w, torch.zeros(c, device=dev), cin_split=c if cat else None), x, x2))
for conv, x, x2 in cases:
    conv(x, x2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for conv, x, x2 in cases:
    conv(x, x2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
