"""GPU parity of the on-device metrics against a restatement of utilities/metrics.py (same loops, same nanmean)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_metrics(logits, mask, n_classes, smooth=1e-10):
    """utilities/metrics.py:8-61 restated (argmax of softmax, classes 1..n-1, absent labels -> nan, nanmean)."""
    pred = torch.argmax(F.softmax(logits, dim=1), dim=1).view(-1)
    m = mask.view(-1)
    acc = float(torch.eq(pred, m).int().sum()) / float(m.numel())
    ious, dices = [], []
    for c in range(1, n_classes):
        tc, tl = pred == c, m == c
        if tl.long().sum().item() == 0:
            ious.append(np.nan); dices.append(np.nan)
            continue
        inter = torch.logical_and(tc, tl).sum().float().item()
        union = torch.logical_or(tc, tl).sum().float().item()
        ious.append((inter + smooth) / (union + smooth))
        dices.append(2 * (inter + smooth) / (union + inter + smooth))
    return acc, np.nanmean(ious), np.nanmean(dices)


@pytest.mark.parametrize("C,B,H,W,drop", [(4, 4, 64, 64, None), (2, 3, 32, 48, None), (4, 2, 64, 64, 2), (7, 2, 33, 35, 5)])
def test_metrics_match_reference(C, B, H, W, drop):
    from uaps_b200 import metrics as M
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C + H)
    logits = (torch.randn(B, C, H, W, generator=g) * 2).to(dev)
    mask = torch.randint(0, C, (B, H, W), generator=g).to(dev)
    if drop is not None:
        mask[mask == drop] = 0                      # a class absent from the labels -> skipped by nanmean
    logits[:, 1] = torch.where(torch.rand(B, H, W, generator=g).to(dev) < 0.1, logits[:, 0], logits[:, 1])   # exact ties
    acc, miou, mdice = _ref_metrics(logits, mask, C)
    assert M.pixel_accuracy(logits, mask) == pytest.approx(acc, rel=1e-12)
    assert M.mIoU(logits, mask, n_classes=C) == pytest.approx(miou, rel=1e-9)
    assert M.mDice(logits, mask, n_classes=C) == pytest.approx(mdice, rel=1e-9)
    conf = M.confusion(logits, mask)
    assert int(conf.sum()) == B * H * W
    accu = M.MetricAccumulator(dev)
    accu.update(logits, mask); accu.update(logits, mask)
    r = accu.result()
    assert r["mIoU"] == pytest.approx(miou, rel=1e-9) and r["pixel_accuracy"] == pytest.approx(acc, rel=1e-12)


def test_predict_fast_path_and_checkpoint_roundtrip(tmp_path):
    from uaps_b200.unet import UNet_UAPS, load_checkpoint, save_checkpoint
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = UNet_UAPS(3, 4, compute="fp32").to(dev).eval()
    x = torch.randn(2, 3, 64, 64, device=dev)
    with torch.no_grad():
        full = m(x)[0]
    assert torch.allclose(m.predict(x), full, rtol=1e-5, atol=1e-5)      # eval-mode BN: main decoder is deterministic
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    path = str(tmp_path / "ck.pth")
    save_checkpoint(path, m, opt, epoch=7, best_dice=0.85)
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"epoch", "best_dice_1", "state_dict", "optimizer"}
    assert all(k.startswith("module.") for k in ck["state_dict"]) and len(ck["state_dict"]) == 334
    m2 = UNet_UAPS(3, 4, compute="fp32").to(dev).eval()
    assert load_checkpoint(path, m2, map_location=dev) == (7, 0.85)
    assert torch.equal(m2.predict(x), m.predict(x))
    m16 = UNet_UAPS(3, 4, compute="bf16").to(dev).eval()
    load_checkpoint(path, m16, map_location=dev)
    p16 = m16.predict(x)
    assert p16.shape == full.shape and ((p16 - full).norm() / full.norm()).item() < 0.1


def test_metrics_match_reference_golden():
    """tests/golden/metrics.npz holds the outputs of the reference's OWN utilities/metrics.py functions
    (oracle/make_golden.py:golden_metrics) -- the on-device confusion-matrix metrics must reproduce them."""
    from conftest import load_cases
    from uaps_b200 import metrics as M
    dev = torch.device("cuda:0")
    for name, c in load_cases("metrics.npz").items():
        logits, mask, C = torch.from_numpy(c["logits"]).to(dev), torch.from_numpy(c["mask"]).to(dev), int(c["n_classes"])
        assert M.pixel_accuracy(logits, mask) == pytest.approx(float(c["pixel_accuracy"]), rel=1e-12), name
        assert M.mIoU(logits, mask, n_classes=C) == pytest.approx(float(c["mIoU"]), rel=1e-9), name
        assert M.mDice(logits, mask, n_classes=C) == pytest.approx(float(c["mDice"]), rel=1e-9), name
        acc = M.MetricAccumulator(dev)                    # per-batch means, as the reference's loop averages them
        acc.update(logits, mask)
        assert acc.result()["mDice"] == pytest.approx(float(c["mDice"]), rel=1e-9), name


def test_out_of_range_labels_count_as_wrong_like_the_reference():
    """utilities/metrics.py:8-13 divides the number of correct pixels by mask.numel(): a label outside [0, C) can never be
    correct but stays in the denominator."""
    from uaps_b200 import metrics as M
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn(2, 4, 32, 32, generator=g, device=dev)
    mask = torch.randint(0, 4, (2, 32, 32), generator=g, device=dev)
    mask[0, :8] = 255                                       # an "ignore" region
    pred = torch.argmax(torch.softmax(logits, 1), 1)
    want = float(torch.eq(pred, mask).sum()) / float(mask.numel())
    assert M.pixel_accuracy(logits, mask) == pytest.approx(want, rel=1e-12)
    acc = M.MetricAccumulator(dev)
    acc.update(logits, mask)
    assert acc.result()["pixel_accuracy"] == pytest.approx(want, rel=1e-12)
