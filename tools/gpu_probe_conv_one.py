import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.conv import PackedConv, to_nhwc_bf16
dev = "cuda:0"
for (B, H, W, ci, co, split) in [(64, 32, 32, 256, 128, 128), (64, 256, 256, 16, 16, None), (64, 64, 64, 64, 64, None)]:
    x = torch.randn(B, ci, H, W, device=dev)
    w = torch.randn(co, ci, 3, 3, device=dev) * 0.05
    conv = PackedConv(w, torch.randn(co, device=dev), cin_split=split)
    xs = (to_nhwc_bf16(x),) if split is None else (to_nhwc_bf16(x[:, :split]), to_nhwc_bf16(x[:, split:]))
    for _ in range(3):
        conv(*xs)
torch.cuda.synchronize()
