"""GPU: one full UAPS iteration through UAPSTrainer against the oracle's iteration (functional U-Net +
restated losses + autograd + Adam) with every random draw injected."""
import math

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
    model = UNet_UAPS(3, C, compute="fp32")            # reference-precision path: the oracle is fp32
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


# ---- device-resident iteration state and the captured (CUDA graph) iteration -------------------------------------------
def test_step_begin_draws_ramp_and_adam_corrections_on_device():
    """uaps_step_begin replaces the host work of UAPS_train.py:251 (Dirichlet), UAPS_unet.py:165 (u ~ U(0.7, 0.9)),
    :279-280 (sigmoid ramp) and Adam's bias correction: check each against the host formula it replaces."""
    import math
    from uaps_b200.ramps import get_current_consistency_weight
    from uaps_b200.stepctx import DeviceStepState
    dev = torch.device("cuda:0")
    a, b = DeviceStepState(dev, 1e-3), DeviceStepState(dev, 1e-3)
    a.set("iter", 7990); b.set("iter", 7990)
    seen = []
    for it in range(7990, 7994):
        a.begin(111, 999, 4, 16, 0.1, 0.07, 200.0, 80, 2, 0.9, 0.999)
        b.begin(222, 999, 4, 16, 0.1, 0.07, 200.0, 80, 2, 0.9, 0.999)        # another rank: same shared seed
        sa, sb = a.read(), b.read()
        assert sa.iter == it + 1
        w = list(sa.mix_w)
        assert all(x > 0 for x in w[:4]) and w[4:] == [0.0] * 4 and sum(w) == pytest.approx(1.0, abs=1e-6)
        assert list(sb.mix_w) == w and list(sb.u) == list(sa.u)           # what every rank must agree on
        assert sa.key_rank != sb.key_rank and sa.key_shared == sb.key_shared
        assert all(0.7 <= u <= 0.9 for u in sa.u)
        assert sa.cw1 == pytest.approx(get_current_consistency_weight(it, 0.1, 200.0, 80), rel=1e-6)
        assert sa.cw2 == pytest.approx(get_current_consistency_weight(it, 0.07, 200.0, 80), rel=1e-6)
        t = it - 7990 + 1
        assert sa.adam_step == t
        assert sa.adam_step_size == pytest.approx(1e-3 / (1 - 0.9 ** t), rel=1e-6)
        assert sa.adam_inv_bc2_sqrt == pytest.approx(1 / math.sqrt(1 - 0.999 ** t), rel=2e-5)      # beta2 crosses the ABI as fp32
        assert (sa.xchg_base, sa.xchg_next) == (2 * (t - 1), 2 * t)
        seen.append((tuple(w), sa.key_rank))
    assert len(set(seen)) == 4                                            # fresh draws every iteration
    ws = np.array([s[0][:4] for s in seen])
    assert ws.std() > 0.01


def _fixed_batch(dev, B=4, HW=64, C=4, seed=4):
    g = torch.Generator().manual_seed(seed)
    xl, xu = torch.randn(B, 3, HW, HW, generator=g).to(dev), torch.randn(B, 3, HW, HW, generator=g).to(dev)
    yl = ((xl[:, 0] > 0).long() + 2 * (xl[:, 1] > 0).long()) % C
    return xl, yl, xu


def test_captured_iteration_replays_the_device_state_iteration():
    """The CUDA-graph replay runs the SAME launch sequence as the eager device-state iteration: from identical weights,
    seeds and data the two loss curves agree step for step (to the run-to-run noise of the fp32 atomics in the weight
    gradient / BatchNorm sums), the model learns, and the graph is captured once."""
    from uaps_b200.train import UAPSConfig, UAPSTrainer
    from uaps_b200.unet import UNet_UAPS
    dev = torch.device("cuda:0")
    xl, yl, xu = _fixed_batch(dev)
    curves = {}
    for graph in (False, True):
        torch.manual_seed(0)
        model = UNet_UAPS(3, 4).to(dev)
        tr = UAPSTrainer(model, UAPSConfig(cuda_graph=graph, graph_warmup=2))
        assert tr.state is not None, "bf16 + FlatAdam must run in device-state mode"
        curves[graph] = [float(tr.step(xl, yl, xu)["loss"]) for _ in range(12)]
        assert tr.iter_num == 12 and tr.state.read().iter == 12 and tr.optimizer.step_count == 12
        assert len(tr._graphs) == (1 if graph else 0)
        assert tr.skipped_steps() == 0
    a, b = curves[False], curves[True]
    assert all(math.isfinite(v) for v in a + b)
    assert b[-1] < 0.9 * b[0], b
    for s, (u, v) in enumerate(zip(a, b)):
        assert abs(u - v) <= 0.03 * abs(u), (s, a, b)


def test_captured_iteration_new_shape_new_graph_and_lr_change():
    from uaps_b200.train import UAPSConfig, UAPSTrainer
    from uaps_b200.unet import UNet_UAPS
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    model = UNet_UAPS(3, 4).to(dev)
    tr = UAPSTrainer(model, UAPSConfig(graph_warmup=1))
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(tr.optimizer, mode="max", factor=0.5, patience=0)   # UAPS_train.py:113
    b1, b2 = _fixed_batch(dev, B=2, HW=64), _fixed_batch(dev, B=2, HW=96)
    for _ in range(3):
        tr.step(*b1)
    for _ in range(3):
        tr.step(*b2)
    assert len(tr._graphs) == 2
    sched.step(1.0); sched.step(0.5)                                        # no improvement -> lr halves
    assert tr.optimizer.param_groups[0]["lr"] == pytest.approx(5e-4)
    before = tr.optimizer.flat_p.clone()
    out = tr.step(*b1)
    st = tr.state.read()
    assert st.lr == pytest.approx(5e-4) and st.adam_step == 7
    assert math.isfinite(float(out["loss"])) and not torch.equal(before, tr.optimizer.flat_p)


def test_non_finite_loss_skips_the_update_on_device():
    """a NaN loss (an exchange that timed out) must not reach the parameters or the moments."""
    from uaps_b200.stepctx import DeviceStepState
    from uaps_b200.train import FlatAdam, FlatGradBuffer
    dev = torch.device("cuda:0")
    ps = [torch.nn.Parameter(torch.randn(33, device=dev)), torch.nn.Parameter(torch.randn(4, 5, device=dev))]
    buf = FlatGradBuffer(ps)
    opt = FlatAdam(buf, lr=1e-2)
    st = DeviceStepState(dev, 1e-2)
    opt.attach_state(st)
    buf.flat.normal_()
    good, bad = torch.tensor(1.5, device=dev), torch.tensor(float("nan"), device=dev)
    st.begin(1, 2, 4, 16, 0.1, 0.1, 200.0, 80, 0, 0.9, 0.999)
    opt.step(guard=good, use_device_state=True)
    p1, m1 = opt.flat_p.clone(), opt.exp_avg.clone()
    st.begin(1, 2, 4, 16, 0.1, 0.1, 200.0, 80, 0, 0.9, 0.999)
    opt.step(guard=bad, use_device_state=True)
    assert torch.equal(opt.flat_p, p1) and torch.equal(opt.exp_avg, m1)
    s = st.read()
    assert (s.skipped, s.n_skipped, s.adam_step) == (1, 1, 2)
    st.begin(1, 2, 4, 16, 0.1, 0.1, 200.0, 80, 0, 0.9, 0.999)             # the skipped step is not consumed
    assert st.read().adam_step == 2 and st.read().skipped == 0
    opt.step(guard=good, use_device_state=True)
    assert not torch.equal(opt.flat_p, p1)


@pytest.mark.parametrize("shape", [(2, 3, 240, 640, 2, 4), (2, 1, 128, 128, 2, 3)])
def test_k5_and_dataset_shapes_run_on_the_kernel_path(shape):
    """BASELINE configs[3]/[4]: KoSDD2-shaped 240x640 with FIVE decoders (4 aux: the 4th re-uses FeatureNoise with a fresh
    draw, through the same perturb3 kernel), DAGM-shaped grayscale C=2.  One captured iteration must train."""
    from uaps_b200.train import UAPSConfig, UAPSTrainer
    from uaps_b200.unet import UNet_UAPS
    B, cin, H, W, C, n_aux = shape
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    model = UNet_UAPS(cin, C, n_aux=n_aux).to(dev)
    tr = UAPSTrainer(model, UAPSConfig(num_classes=C, graph_warmup=1))
    assert tr.k == n_aux + 1
    g = torch.Generator().manual_seed(0)
    xl, xu = torch.randn(B, cin, H, W, generator=g).to(dev), torch.randn(B, cin, H, W, generator=g).to(dev)
    yl = (xl[:, 0] > 0).long()
    losses = [float(tr.step(xl, yl, xu)["loss"]) for _ in range(6)]
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0], losses
    outs = model(xu)
    assert len(outs) == n_aux + 1 and all(o.shape == (B, C, H, W) and o.dtype == torch.float32 for o in outs)


def test_step_host_prefetch_matches_step():
    """step_host (pinned host batch, H2D on a side stream into double-buffered staging) runs the same iteration as step."""
    from uaps_b200.train import UAPSConfig, UAPSTrainer
    from uaps_b200.unet import UNet_UAPS
    dev = torch.device("cuda:0")
    batches = [tuple(t.cpu().pin_memory() for t in _fixed_batch(dev, B=2, HW=64, seed=s)) for s in range(3)]
    curves = {}
    for host in (False, True):
        torch.manual_seed(0)
        tr = UAPSTrainer(UNet_UAPS(3, 4).to(dev), UAPSConfig(graph_warmup=1))
        out = []
        for i in range(9):
            b = batches[i % 3]
            o = tr.step_host(*b) if host else tr.step(*(t.to(dev) for t in b))
            out.append(float(o["loss"]))
        curves[host] = out
    for s, (u, v) in enumerate(zip(curves[False], curves[True])):
        assert math.isfinite(v) and abs(u - v) <= 0.03 * abs(u), (s, curves)
