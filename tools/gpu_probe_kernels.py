"""One launch of every hand-written kernel of the bf16 training path at the level-0 shape (B x 256 x 256, 16 channels),
between cudaProfilerStart/Stop, for `ncu --set full --profile-from-start off` (profiles/r02_kernels_*.txt)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200.bn_act import bn_lrelu_dropout
from uaps_b200.conv import PackedConv, conv_wgrad, to_nhwc_bf16
from uaps_b200.perturb import perturb3_nhwc
from uaps_b200.resample import maxpool2, upsample2x

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
which = sys.argv[2] if len(sys.argv) > 2 else "all"
H = W = 256
torch.manual_seed(0)


def cl(c, h=H, w=W, grad=False):
    t = torch.randn(B, h, w, c, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
    return t.requires_grad_(grad)


def run(warm):
    out = {}
    if which in ("all", "conv"):
        x16 = cl(16).permute(0, 2, 3, 1).contiguous()
        x16b = cl(16).permute(0, 2, 3, 1).contiguous()
        w = torch.randn(16, 16, 3, 3, device=dev) * 0.05
        wc = torch.randn(16, 32, 3, 3, device=dev) * 0.05
        wo = torch.randn(4, 16, 3, 3, device=dev) * 0.05
        b16, b4 = torch.zeros(16, device=dev), torch.zeros(4, device=dev)
        c_f, c_cat, c_out = PackedConv(w, b16), PackedConv(wc, b16, cin_split=16), PackedConv(wo, b4)
        c_d, c_dcat = PackedConv(w, None, transpose=True), PackedConv(wc, None, transpose=True)
        dw = torch.zeros(16, 16, 3, 3, device=dev)
        c_f(x16); c_cat(x16, x16b); c_out(x16, out_nchw_f32=True); c_d(x16); c_dcat(x16, split=16)
        conv_wgrad(x16, [x16b], 16, 16, 3, out=dw)
        # tensor-bound layers: up1.c1 cat 128+128 -> 128 @ 32x32 and enc3.c2 128 -> 128
        xa, xb = (torch.randn(B, 32, 32, 128, device=dev).to(torch.bfloat16) for _ in range(2))
        w128 = torch.randn(128, 256, 3, 3, device=dev) * 0.02
        PackedConv(w128, torch.zeros(128, device=dev), cin_split=128)(xa, xb)
        dw128 = torch.zeros(128, 128, 3, 3, device=dev)
        conv_wgrad(xa, [xb], 128, 128, 3, out=dw128)
    if which in ("all", "elem"):
        bn = torch.nn.BatchNorm2d(16).to(dev)
        y = cl(16, grad=True)
        a = bn_lrelu_dropout(y, bn, 0.05)
        a.backward(torch.randn_like(a))
        f = cl(16, grad=True)
        ys = perturb3_nhwc(f)
        (ys[0].float().sum() * 0 + sum((t * 1.0).sum() for t in ys)).backward() if False else torch.autograd.backward(ys, [torch.ones_like(t) for t in ys])
        xs = cl(16, 128, 128, grad=True)
        up = upsample2x(xs)
        up.backward(torch.randn_like(up))
        xp = cl(16, grad=True)
        mp = maxpool2(xp)
        mp.backward(torch.randn_like(mp))
        to_nhwc_bf16(torch.randn(B, 3, H, W, device=dev))
    torch.cuda.synchronize()


run(True)
torch.cuda.profiler.start()
run(False)
torch.cuda.profiler.stop()
print("done")
