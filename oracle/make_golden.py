"""Pin the oracle to the real reference and write tests/golden/*.npz (TEST INFRASTRUCTURE).

Runs ONLY in the build container, where /root/reference exists:

    python -m oracle.make_golden

1. imports the reference's own modules (utilities/UAPS_unet.py, pytorch_losses.py, ramps.py)
   and *executes the reference's inline loss lines* (UAPS_train.py:186-189, 194-218, 223-277,
   read from the file at run time, never copied into this repo) with the random draws patched
   to injected values;
2. asserts that every oracle restatement reproduces them (bit-exact where the op order is the
   same, 1e-6 otherwise);
3. writes the input/output vectors as small fixtures for the CPU and GPU parity tests.

The GPU box has no /root/reference; there the tests use the fixtures and the oracle only.
"""
from __future__ import annotations

import os
import sys
import textwrap

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)          # NOT /root/reference/utilities: utilities/utilities.py would shadow the package
    import utilities.UAPS_unet as ref_unet
    import utilities.pytorch_losses as ref_losses
    import utilities.ramps as ref_ramps
    return ref_unet, ref_losses, ref_ramps


def _reference_lines(first: int, last: int) -> str:
    with open(os.path.join(REF, "UAPS_train.py")) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


class _FakeNP:
    """Stands in for ``np`` inside the exec'd reference lines: dirichlet returns the injected draw."""

    def __init__(self, w):
        self._w = np.asarray(w, dtype=np.float64)
        self.random = self

    def dirichlet(self, alpha, size=1):
        assert len(alpha) == len(self._w)
        return self._w[None, :]

    def ones(self, n):
        return np.ones(n)


def run_reference_unlabeled(logits, mix_w, cw1, cw2, ref_losses):
    """Execute UAPS_train.py:186-189 + 223-277 on K=4 logits tensors."""
    import torch.nn as nn
    ns = {"torch": torch, "np": _FakeNP(mix_w),
          "kl_distance": nn.KLDivLoss(reduction="none"), "log_sm": nn.LogSoftmax(dim=1),
          "ce_loss": nn.CrossEntropyLoss(), "dice_loss": ref_losses.dice_loss}
    ns["un_outputs"], ns["un_outputs_aux1"], ns["un_outputs_aux2"], ns["un_outputs_aux3"] = logits
    exec(_reference_lines(186, 189), ns)
    exec(_reference_lines(223, 277), ns)
    loss_u = cw1 * ns["ps_loss"] + cw2 * ns["l_uncert"]
    return {"loss_u": loss_u, "ps_loss": ns["ps_loss"], "l_uncert": ns["l_uncert"],
            "pseudo": ns["un_lbl_pseudo"],
            "exp_var": [ns["exp_variance_main"], ns["exp_variance_aux1"], ns["exp_variance_aux2"],
                        ns["exp_variance_aux3"]]}


def run_reference_supervised(logits, labels, ref_losses):
    import torch.nn as nn
    ns = {"torch": torch, "ce_loss": nn.CrossEntropyLoss(), "dice_loss": ref_losses.dice_loss,
          "labels_S1": labels}
    ns["outputs"], ns["outputs_aux1"], ns["outputs_aux2"], ns["outputs_aux3"] = logits
    exec(_reference_lines(194, 218), ns)
    return {"supervised_loss": ns["supervised_loss"], "total_loss_ce": ns["total_loss_ce"],
            "total_loss_dice": ns["total_loss_dice"]}


def _logits(K, B, C, H, W, seed, scale=2.0, ties=False):
    out = []
    for k in range(K):
        g = torch.Generator().manual_seed(seed + k)
        z = torch.randn(B, C, H, W, generator=g) * scale
        if ties:                        # duplicate class planes -> exact ties in the mixed prediction
            z[:, 1] = z[:, 0]
            if C > 3:
                z[:, 3] = z[:, 2]
        out.append(z)
    return out


def golden_losses(ref_losses):
    from oracle.uaps_loss_ref import (dice_loss_ref, supervised_loss_ref, unlabeled_loss_ref,
                                      unlabeled_loss_fp64_closed_form)
    cases = [  # name, K, B, C, H, W, scale, ties
        ("k4c4", 4, 2, 4, 8, 16, 2.0, False),
        ("k4c2", 4, 3, 2, 4, 12, 2.0, False),
        ("k4c4_peaked", 4, 2, 4, 8, 8, 16.0, False),
        ("k4c4_ties", 4, 2, 4, 8, 8, 2.0, True),
        ("k4c3_odd", 4, 1, 3, 5, 7, 2.0, False),     # HW = 35: exercises the unaligned path
        ("k2c2", 2, 2, 2, 8, 8, 2.0, False),         # K != 4: restatement only (reference is written for K = 4)
        ("k3c4", 3, 2, 4, 8, 8, 2.0, False),
        ("k5c2", 5, 2, 2, 8, 8, 2.0, False),
        ("k6c4", 6, 1, 4, 8, 8, 2.0, False),
    ]
    store = {}
    for name, K, B, C, H, W, scale, ties in cases:
        mix_w = np.random.default_rng(1337).dirichlet(np.ones(K))
        cw1, cw2 = 0.1, 0.07
        z = [t.requires_grad_(True) for t in _logits(K, B, C, H, W, 1337, scale, ties)]
        mine = unlabeled_loss_ref(z, mix_w, cw1, cw2)
        mine["loss_u"].backward()
        grads = [t.grad.clone() for t in z]
        if K == 4:
            z2 = [t.detach().clone().requires_grad_(True) for t in z]
            ref = run_reference_unlabeled(z2, mix_w, cw1, cw2, ref_losses)
            ref["loss_u"].backward()
            assert torch.equal(ref["pseudo"], mine["pseudo"]), name
            for a, b in (("loss_u",) * 2, ("ps_loss",) * 2, ("l_uncert",) * 2):
                assert torch.equal(ref[a], mine[b]), (name, a, ref[a].item(), mine[b].item())
            for k in range(K):
                assert torch.equal(ref["exp_var"][k], mine["exp_var"][k]), name
                assert torch.equal(z2[k].grad, grads[k]), name
            pinned = 1
        else:
            pinned = 0
        # fp64 closed form agrees with autograd of the verbatim expressions
        cf = unlabeled_loss_fp64_closed_form(z, mix_w, cw1, cw2, mine["pseudo"])
        for k in range(K):
            scale_g = grads[k].abs().max().item()
            err = (cf["dz"][k].float() - grads[k]).abs().max().item()
            assert err <= 2e-5 * scale_g + 1e-12, (name, k, err, scale_g)
        store[name] = dict(
            logits=torch.stack([t.detach() for t in z]).numpy(), mix_w=mix_w, cw=np.array([cw1, cw2]),
            loss_u=mine["loss_u"].item(), ps_loss=mine["ps_loss"].item(), l_uncert=mine["l_uncert"].item(),
            pseudo=mine["pseudo"].numpy().astype(np.int64),
            exp_var=torch.stack([e.detach() for e in mine["exp_var"]]).numpy(),
            ps=np.array([p.item() for p in mine["ps"]]),
            grads=torch.stack(grads).numpy(), pinned_to_reference_lines=pinned)
        print(f"  loss case {name}: pinned={pinned} loss_u={mine['loss_u'].item():.6f}")

    # supervised section + dice_loss
    K, B, C, H, W = 4, 2, 4, 8, 16
    z = [t.requires_grad_(True) for t in _logits(K, B, C, H, W, 4242)]
    labels = torch.randint(0, C, (B, H, W), generator=torch.Generator().manual_seed(5))
    mine = supervised_loss_ref(z, labels)
    mine["supervised_loss"].backward()
    z2 = [t.detach().clone().requires_grad_(True) for t in z]
    ref = run_reference_supervised(z2, labels, ref_losses)
    ref["supervised_loss"].backward()
    for key in ("supervised_loss", "total_loss_ce", "total_loss_dice"):
        assert torch.equal(ref[key], mine[key]), key
    for k in range(K):
        assert torch.equal(z2[k].grad, z[k].grad)
        assert torch.equal(ref_losses.dice_loss(labels.unsqueeze(1), z[k].detach()),
                           dice_loss_ref(labels.unsqueeze(1), z[k].detach()))
    store["sup_k4c4"] = dict(
        logits=torch.stack([t.detach() for t in z]).numpy(), labels=labels.numpy(),
        supervised_loss=mine["supervised_loss"].item(), total_loss_ce=mine["total_loss_ce"].item(),
        total_loss_dice=mine["total_loss_dice"].item(),
        ce=np.array([c.item() for c in mine["ce"]]), dice=np.array([d.item() for d in mine["dice"]]),
        grads=torch.stack([t.grad for t in z]).numpy())
    print("  supervised section: pinned")
    flat = {f"{case}/{k}": v for case, d in store.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, "loss_cases.npz"), **flat)


def golden_ramps(ref_ramps):
    from oracle.uaps_loss_ref import consistency_weight_ref, sigmoid_rampup_ref
    cur = np.array([0, 1, 5, 50, 100, 199, 200, 250, -3], dtype=np.float64)
    vals = []
    for c in cur:
        a, b = ref_ramps.sigmoid_rampup(c, 200.0), sigmoid_rampup_ref(c, 200.0)
        assert abs(a - b) <= 1e-15 * max(1.0, abs(a)), (c, a, b)
        vals.append(a)
    assert ref_ramps.sigmoid_rampup(3, 0) == sigmoid_rampup_ref(3, 0) == 1.0
    iters = np.array([0, 79, 80, 1000, 15999, 16000, 47200])
    cw = [0.1 * ref_ramps.sigmoid_rampup(i // 80, 200.0) for i in iters]
    for i, w in zip(iters, cw):
        assert abs(consistency_weight_ref(int(i)) - w) < 1e-15
    np.savez(os.path.join(OUT, "ramps.npz"), current=cur, sigmoid_rampup_200=np.array(vals),
             iters=iters, consistency_weight=np.array(cw))
    print("  ramps: pinned")


def golden_perturb(ref_unet):
    from oracle.perturb_ref import dropout_ref, feature_dropout_ref, feature_noise_ref
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 16, 12, 20, generator=g)
    # FeatureNoise: patch the distribution's sample() to return the injected draw
    noise = (torch.rand(x.shape[1:], generator=g) * 2 - 1) * 0.3
    fn = ref_unet.FeatureNoise()
    fn.uni_dist.sample = lambda shape: noise.clone()
    y_noise = fn(x)
    assert torch.equal(y_noise, feature_noise_ref(x, noise))
    # Dropout: run the reference under a known torch seed, recover the keep mask from the output
    torch.manual_seed(11)
    y_drop = ref_unet.Dropout(x)
    keep = y_drop != 0
    assert torch.equal(y_drop, dropout_ref(x, keep, 0.5))
    assert 0.4 < keep.float().mean().item() < 0.6
    # FeatureDropout: numpy global generator seeded, the same draw replayed for the oracle
    np.random.seed(5)
    y_fd = ref_unet.FeatureDropout(x)
    np.random.seed(5)
    u = float(np.random.uniform(0.7, 0.9))
    assert torch.equal(y_fd, feature_dropout_ref(x, u))
    np.savez_compressed(os.path.join(OUT, "perturb.npz"), x=x.numpy(), noise=noise.numpy(),
                        y_noise=y_noise.numpy(), keep=keep.numpy(), y_drop=y_drop.numpy(),
                        u=np.array(u), y_fd=y_fd.numpy())
    print("  perturbations: pinned")


def golden_unet(ref_unet):
    """Reference UNet_UAPS (train mode) with patched randomness vs the functional oracle."""
    import torch.nn.functional as F
    from torch.distributions.uniform import Uniform
    from oracle.unet_ref import feature_shapes, synthetic_rand, synthetic_state_dict, unet_uaps_ref

    in_chns, C, B, H, W = 3, 4, 2, 32, 32
    sd = synthetic_state_dict(in_chns, C)
    model = ref_unet.UNet_UAPS(in_chns, C)
    missing = set(model.state_dict()) ^ set(sd)
    assert not missing, missing
    assert len(sd) == 334
    model.load_state_dict(sd)
    model.train()
    x = torch.randn(B, in_chns, H, W, generator=torch.Generator().manual_seed(3))
    rand = synthetic_rand(feature_shapes(B, H, W))

    drop_q = list(rand["enc_keep"]) + list(rand["aux2_keep"])
    noise_q = list(rand["noise"])
    u_q = list(rand["u"])
    orig_dropout, orig_sample, orig_uniform = F.dropout, Uniform.sample, np.random.uniform

    def fake_dropout(inp, p=0.5, training=True, inplace=False):
        if p == 0.0 or not training:
            return inp
        return inp * drop_q.pop(0).to(inp.dtype) * (1.0 / (1.0 - p))

    F.dropout = fake_dropout
    Uniform.sample = lambda self, shape=torch.Size(): noise_q.pop(0).clone()
    np.random.uniform = lambda a, b: u_q.pop(0)
    try:
        ref_out = model(x)
    finally:
        F.dropout, Uniform.sample, np.random.uniform = orig_dropout, orig_sample, orig_uniform
    assert not drop_q and not noise_q and not u_q

    sd_req = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
              for k, v in sd.items()}
    stats = {}
    mine = unet_uaps_ref(x, sd_req, rand, stats)
    for a, b in zip(ref_out, mine[:4]):
        err = (a - b).abs().max().item()
        assert err <= 1e-5 * a.abs().max().item(), err
    new_sd = model.state_dict()
    for k, v in stats.items():
        assert torch.allclose(new_sd[k], v, rtol=1e-5, atol=1e-6), k

    # backward golden: fixed cotangents on the four logits tensors
    gcot = torch.Generator().manual_seed(17)
    cot = [torch.randn(o.shape, generator=gcot) for o in ref_out]
    sum((o * c).sum() for o, c in zip(ref_out, cot)).backward()
    sum((o * c).sum() for o, c in zip(mine[:4], cot)).backward()
    names = [n for n, _ in model.named_parameters()]
    gsum, gnorm = [], []
    for n, p in model.named_parameters():
        g_ref, g_mine = p.grad, sd_req[n].grad
        assert torch.allclose(g_ref, g_mine, rtol=2e-3, atol=2e-4 * g_ref.abs().max().item() + 1e-7), n
        gsum.append(g_ref.sum().item())
        gnorm.append(g_ref.norm().item())
    keep_full = ["encoder.in_conv.conv_conv.0.weight", "encoder.down2.maxpool_conv.1.conv_conv.4.weight",
                 "main_decoder.out_conv.weight", "aux_decoder1.up1.conv1x1.weight",
                 "aux_decoder3.up4.conv.conv_conv.1.weight", "aux_decoder2.up2.conv.conv_conv.5.bias"]
    pg = dict(model.named_parameters())
    np.savez_compressed(
        os.path.join(OUT, "unet_small.npz"), x=x.numpy(),
        out=torch.stack([o.detach() for o in ref_out]).numpy(), cot=torch.stack(cot).numpy(),
        feat_sum=np.array([f.detach().double().sum().item() for f in mine[4]]),
        grad_names=np.array(names), grad_sum=np.array(gsum), grad_norm=np.array(gnorm),
        **{"grad/" + n: pg[n].grad.numpy() for n in keep_full},
        **{"stat/" + k: v.numpy() for k, v in stats.items() if k.startswith("encoder.in_conv")})
    print(f"  UNet_UAPS forward/backward (B={B}, {H}x{W}): pinned, 334 state tensors")


def golden_metrics():
    """utilities/metrics.py:7-61 (pixel_accuracy, mIoU, mDice) evaluated by the reference's own functions on seeded
    logits / masks, including an absent class (nanmean path), exact logit ties (argmax tie-break) and C != 4."""
    import utilities.metrics as ref_metrics
    out = {}
    for name, (C, B, H, W, drop) in {"c4": (4, 4, 64, 64, None), "c2": (2, 3, 32, 48, None), "c4_absent2": (4, 2, 64, 64, 2),
                                     "c7_absent5": (7, 2, 33, 35, 5)}.items():
        g = torch.Generator().manual_seed(100 + C + H)
        logits = torch.randn(B, C, H, W, generator=g) * 2
        mask = torch.randint(0, C, (B, H, W), generator=g)
        if drop is not None:
            mask[mask == drop] = 0
        logits[:, 1] = torch.where(torch.rand(B, H, W, generator=g) < 0.1, logits[:, 0], logits[:, 1])     # exact ties
        out[name + "/logits"] = logits.numpy()
        out[name + "/mask"] = mask.numpy()
        out[name + "/n_classes"] = np.array(C)
        out[name + "/pixel_accuracy"] = np.array(ref_metrics.pixel_accuracy(logits, mask), dtype=np.float64)
        out[name + "/mIoU"] = np.array(ref_metrics.mIoU(logits, mask, n_classes=C), dtype=np.float64)
        out[name + "/mDice"] = np.array(ref_metrics.mDice(logits, mask, n_classes=C), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)
    print(f"  metrics (pixel_accuracy / mIoU / mDice): {len(out) // 6} cases from the reference's functions")


def main():
    if not os.path.isdir(REF):
        raise SystemExit("make_golden needs /root/reference (build container only)")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    ref_unet, ref_losses, ref_ramps = _import_reference()
    print("pinning oracle to reference modules:")
    golden_ramps(ref_ramps)
    golden_perturb(ref_unet)
    golden_losses(ref_losses)
    golden_unet(ref_unet)
    golden_metrics()
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
