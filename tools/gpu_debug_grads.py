import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import uaps_b200.unet as U
from uaps_b200.unet import UNet_UAPS
from oracle.unet_ref import feature_shapes, synthetic_rand, synthetic_state_dict
from test_unet_gpu import _torch_bf16_conv, _to
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
sd = synthetic_state_dict(3, 4, seed=3)
m32, m16, mref = UNet_UAPS(3, 4), UNet_UAPS(3, 4, compute="bf16"), UNet_UAPS(3, 4, compute="bf16")
for m in (m32, m16, mref):
    m.load_state_dict(sd); m.to(dev).train()
x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(1)).to(dev)
rand = _to(synthetic_rand(feature_shapes(2, 128, 128), seed=9), dev)
o32, o16 = m32(x, rand=rand), m16(x, rand=rand)
ours = U.conv_bf16
U.conv_bf16 = _torch_bf16_conv
oref = mref(x, rand=rand)
cot = [torch.randn(o.shape, generator=torch.Generator().manual_seed(k)).to(dev) for k, o in enumerate(o32)]
sum((o * c).sum() for o, c in zip(oref, cot)).backward()
U.conv_bf16 = ours
sum((o * c).sum() for o, c in zip(o32, cot)).backward()
sum((o * c).sum() for o, c in zip(o16, cot)).backward()
g32, g16, gref = dict(m32.named_parameters()), dict(m16.named_parameters()), dict(mref.named_parameters())
rows = []
for n in g32:
    if n.endswith(('conv_conv.0.bias', 'conv_conv.4.bias')): continue
    a, r, b = g32[n].grad.flatten().double(), gref[n].grad.flatten().double(), g16[n].grad.flatten().double()
    cos = lambda u, v: (u @ v / (u.norm() * v.norm() + 1e-30)).item()
    rows.append((cos(r, b), cos(a, b), cos(a, r), a.norm().item(), b.norm().item(), r.norm().item(), n))
rows.sort()
print("cos(cudnn16,ours) cos(fp32,ours) cos(fp32,cudnn16) |g32| |ours| |cudnn16| name")
for r in rows[:25]: print("%.3f %.3f %.3f %.3e %.3e %.3e %s" % r)
print("...")
for r in rows[-5:]: print("%.3f %.3f %.3f %.3e %.3e %.3e %s" % r)

import torch
A = torch.cat([g32[r[-1]].grad.flatten().double() for r in rows]); B = torch.cat([g16[r[-1]].grad.flatten().double() for r in rows]); R = torch.cat([gref[r[-1]].grad.flatten().double() for r in rows])
c = lambda u, v: (u @ v / (u.norm() * v.norm())).item()
print("global: fp32-ours %.4f fp32-cudnn16 %.4f cudnn16-ours %.4f" % (c(A, B), c(A, R), c(R, B)))
