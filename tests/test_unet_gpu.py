"""GPU parity of the drop-in UNet_UAPS against the golden vectors produced by the reference's own
UNet_UAPS (randomness injected) and against the functional oracle on device."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.unet_ref import feature_shapes, synthetic_rand, synthetic_state_dict, unet_uaps_ref

pytestmark = pytest.mark.gpu


def _to(rand, dev):
    return {k: [v.to(dev) if torch.is_tensor(v) else v for v in vals] for k, vals in rand.items()}


def test_state_dict_keys_and_golden_forward_backward():
    from uaps_b200.unet import UNet_UAPS, load_reference_state_dict
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLDEN, "unet_small.npz"))
    sd = synthetic_state_dict(3, 4)
    model = UNet_UAPS(3, 4)
    assert list(model.state_dict().keys()) != [] and set(model.state_dict().keys()) == set(sd.keys())
    assert len(model.state_dict()) == 334
    load_reference_state_dict(model, {"module." + k: v for k, v in sd.items()})     # DataParallel-style keys
    model = model.to(dev).train()
    x = torch.from_numpy(g["x"]).to(dev)
    B, _, H, W = x.shape
    rand = _to(synthetic_rand(feature_shapes(B, H, W)), dev)
    out = model(x, rand=rand)
    ref = torch.from_numpy(g["out"]).to(dev)
    for k in range(4):
        err = (out[k] - ref[k]).abs().max().item()
        assert err <= 2e-4 * ref[k].abs().max().item(), (k, err)
    cot = torch.from_numpy(g["cot"]).to(dev)
    sum((o * c).sum() for o, c in zip(out, cot)).backward()
    grads = dict(model.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    for n, gn in zip(names, g["grad_norm"]):
        # a conv bias that feeds a train-mode BatchNorm has an analytically zero gradient (BN removes
        # the mean); what the reference stores there is rounding noise, so only real gradients are compared
        if n.endswith(("conv_conv.0.bias", "conv_conv.4.bias")):
            assert grads[n].grad.norm().item() <= 1e-2, n
            continue
        # LeakyReLU slope flips at pre-activations within rounding of zero make parameter gradients
        # discontinuous in the forward roundings: cuDNN-vs-CPU fp32 already differ by ~1% here
        assert grads[n].grad.norm().item() == pytest.approx(float(gn), rel=3e-2, abs=1e-5), n
    for key in g.files:
        if key.startswith("grad/"):
            want = torch.from_numpy(g[key]).to(dev)
            got = grads[key[5:]].grad
            assert (got - want).abs().max().item() <= 3e-2 * want.abs().max().item() + 1e-6, key
    # BatchNorm running statistics advanced like the reference's
    new_sd = model.state_dict()
    for key in g.files:
        if key.startswith("stat/"):
            assert torch.allclose(new_sd[key[5:]].cpu(), torch.from_numpy(g[key]), rtol=1e-4, atol=1e-5), key


def test_against_oracle_on_device_neu_shape():
    from uaps_b200.unet import UNet_UAPS
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    sd = synthetic_state_dict(3, 4, seed=7)
    model = UNet_UAPS(3, 4)
    model.load_state_dict(sd)
    model = model.to(dev).train()
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(0)).to(dev)
    rand = _to(synthetic_rand(feature_shapes(2, 256, 256), seed=5), dev)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    ref = unet_uaps_ref(x, sd_dev, rand)
    out = model(x, rand=rand)
    for k in range(4):
        err = (out[k] - ref[k]).abs().max().item()
        assert err <= 1e-4 * ref[k].abs().max().item(), (k, err)


def test_net_factory_and_input_check():
    from uaps_b200.unet import net_factory
    assert net_factory("unet_ccps") is None            # the reference returns None for unknown names
    m = net_factory("unet_uaps", in_chns=1, class_num=2)
    assert next(m.parameters()).is_cuda
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 1, 200, 200, device="cuda"))   # 200 is not a multiple of 16 (reference fails in cat)
    out = m(torch.randn(2, 1, 64, 96, device="cuda"))
    assert len(out) == 4 and out[0].shape == (2, 2, 64, 96)
