"""GPU: one full UAPS iteration through UAPSTrainer against the oracle's iteration (functional U-Net +
restated losses + autograd + Adam) with every random draw injected."""
import numpy as np
import pytest
import torch

from oracle.unet_ref import feature_shapes, synthetic_rand, synthetic_state_dict, unet_uaps_ref
from oracle.uaps_loss_ref import consistency_weight_ref, supervised_loss_ref, unlabeled_loss_ref

pytestmark = pytest.mark.gpu


def test_training_iteration_matches_oracle():
    from uaps_b200.train import UAPSTrainer
    from uaps_b200.unet import UNet_UAPS
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, C = 2, 64, 64, 4
    sd = synthetic_state_dict(3, C, seed=11)
    model = UNet_UAPS(3, C)
    model.load_state_dict(sd)
    model = model.to(dev)
    trainer = UAPSTrainer(model)
    trainer.iter_num = 8000                                   # mid-ramp: cw = 0.1 * exp(-5 (1 - 100/200)^2)
    g = torch.Generator().manual_seed(2)
    xl, xu = torch.randn(B, 3, H, W, generator=g).to(dev), torch.randn(B, 3, H, W, generator=g).to(dev)
    yl = torch.randint(0, C, (B, H, W), generator=g).to(dev)
    to = lambda r: {k: [v.to(dev) if torch.is_tensor(v) else v for v in vals] for k, vals in r.items()}
    rl, ru = to(synthetic_rand(feature_shapes(B, H, W), 1)), to(synthetic_rand(feature_shapes(B, H, W), 2))
    mix_w = np.random.default_rng(3).dirichlet(np.ones(4))

    # oracle iteration on the same device
    params = {k: v.to(dev).clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    full = {**{k: v.to(dev) for k, v in sd.items()}, **params}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    ol, ou = unet_uaps_ref(xl, full, rl)[:4], unet_uaps_ref(xu, full, ru)[:4]
    cw = consistency_weight_ref(8000)
    ref_sup = supervised_loss_ref(ol, yl)
    ref_un = unlabeled_loss_ref(ou, mix_w, cw, cw)
    ref_loss = ref_sup["supervised_loss"] + ref_un["loss_u"]
    opt.zero_grad(); ref_loss.backward(); opt.step()

    out = trainer.step(xl, yl, xu, mix_w=mix_w, rand_l=rl, rand_u=ru)
    assert trainer.consistency_weights()[0] == pytest.approx(consistency_weight_ref(8001), rel=1e-12)
    assert out["loss"].item() == pytest.approx(ref_loss.item(), rel=2e-4)
    assert out["supervised_loss"].item() == pytest.approx(ref_sup["supervised_loss"].item(), rel=2e-4)
    assert out["ps_loss"].item() == pytest.approx(ref_un["ps_loss"].item(), rel=2e-4)
    assert out["l_uncert"].item() == pytest.approx(ref_un["l_uncert"].item(), rel=2e-3)
    # parameters after one Adam step: Adam normalises the gradient, so compare the update direction
    new = dict(model.named_parameters())
    agree, total = 0, 0
    for k, p_ref in params.items():
        if k.endswith(("conv_conv.0.bias", "conv_conv.4.bias")):
            continue                                          # analytically zero gradient in front of train-mode BN
        d_ref, d_new = (p_ref.detach() - sd[k].to(dev)), (new[k].detach() - sd[k].to(dev))
        agree += (torch.sign(d_ref) == torch.sign(d_new)).sum().item()
        total += d_ref.numel()
    assert agree / total > 0.995, agree / total


def test_bf16_path_trains_like_fp32_path():
    """Ten iterations from the same initial weights on the same fixed batch: the bf16 / tcgen05 path's loss
    curve tracks the fp32 path's (within 5% at every step) and both decrease."""
    from uaps_b200 import perturb as P
    from uaps_b200.train import UAPSTrainer
    from uaps_b200.unet import UNet_UAPS
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(4)
    xl, xu = torch.randn(4, 3, 64, 64, generator=g).to(dev), torch.randn(4, 3, 64, 64, generator=g).to(dev)
    yl = (xl[:, 0] > 0).long() + 2 * (xl[:, 1] > 0).long()          # a learnable 4-class target
    curves = {}
    for mode in ("fp32", "bf16"):
        torch.manual_seed(0)
        P.manual_seed(1)
        model = UNet_UAPS(3, 4, compute=mode).to(dev)
        tr = UAPSTrainer(model)
        curves[mode] = [tr.step(xl, yl, xu)["loss"].item() for _ in range(10)]
    a, b = curves["fp32"], curves["bf16"]
    assert a[-1] < 0.9 * a[0] and b[-1] < 0.9 * b[0], (a, b)
    for s, (u, v) in enumerate(zip(a, b)):
        assert abs(u - v) <= 0.05 * abs(u), (s, u, v)


def test_flat_adam_matches_torch_adam_and_shares_its_checkpoint_format():
    """Row f4: FlatAdam (one kernel over flat buffers, uaps_adam_step) against torch.optim.Adam on the same gradients,
    and state_dict round trips in both directions (the reference's checkpoint stores optimizer_1.state_dict(), :447)."""
    from uaps_b200.train import FlatAdam, FlatGradBuffer
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    shapes = [(16, 3, 3, 3), (16,), (7,), (32, 16, 3, 3), (2,), (4, 16, 3, 3)]
    ours = [torch.nn.Parameter(torch.randn(*s, device=dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    buf = FlatGradBuffer(ours)
    opt = FlatAdam(buf, lr=1e-2)
    topt = torch.optim.Adam(ref, lr=1e-2)
    assert all(p.data_ptr() % 256 == 0 and p.grad.data_ptr() % 256 == 0 for p in ours)
    for it in range(7):
        if it == 4:                                          # a scheduler lowers the rate in place (:113)
            opt.param_groups[0]["lr"] = 3e-3
            topt.param_groups[0]["lr"] = 3e-3
        for p, r in zip(ours, ref):
            g = torch.randn_like(r) * (0.1 + it)
            p.grad.copy_(g)
            r.grad = g.clone()
        opt.step(); topt.step()
    for p, r in zip(ours, ref):
        torch.testing.assert_close(p.detach(), r.detach(), rtol=2e-6, atol=2e-7)
    # torch -> ours -> torch: the moments and the step count survive
    sd_t = topt.state_dict()
    opt2 = FlatAdam(FlatGradBuffer([torch.nn.Parameter(p.detach().clone()) for p in ref]), lr=1.0)
    opt2.load_state_dict(sd_t)
    assert opt2.step_count == 7 and opt2.param_groups[0]["lr"] == 3e-3
    sd_o = opt2.state_dict()
    assert set(sd_o) == {"state", "param_groups"} and set(sd_o["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    for i in range(len(shapes)):
        torch.testing.assert_close(sd_o["state"][i]["exp_avg"], sd_t["state"][i]["exp_avg"])
        torch.testing.assert_close(sd_o["state"][i]["exp_avg_sq"], sd_t["state"][i]["exp_avg_sq"])
    topt2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ref], lr=1.0)
    topt2.load_state_dict(sd_o)                              # torch accepts our dict
    assert topt2.param_groups[0]["lr"] == 3e-3
