import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.conv import conv_wgrad, to_nhwc_bf16
dev = "cuda:0"
B = 64
for (H, W, ci, co) in [(256, 256, 16, 16), (128, 128, 32, 32), (64, 64, 64, 64), (32, 32, 128, 128), (16, 16, 256, 256), (256, 256, 32, 16)]:
    x = to_nhwc_bf16(torch.randn(B, ci, H, W, device=dev)); dy = to_nhwc_bf16(torch.randn(B, co, H, W, device=dev))
    for _ in range(2):
        conv_wgrad(dy, [x], co, ci, 3)
torch.cuda.synchronize()
