"""A few launches of the weight-gradient kernel at the shapes that matter (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.conv import conv_wgrad
dev = "cuda:0"
B = 64
for (H, ci, co) in [(256, 16, 16), (128, 32, 32), (32, 128, 128)]:
    x = torch.randn(B, H, H, ci, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, H, H, co, device=dev).to(torch.bfloat16)
    dw = torch.zeros(co, ci, 3, 3, device=dev)
    for _ in range(3):
        conv_wgrad(dy, [x], co, ci, 3, out=dw)
torch.cuda.synchronize()
