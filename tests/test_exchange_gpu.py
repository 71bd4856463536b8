"""The multi-GPU exchange of the fused loss (uaps_loss_pass1_exchange, uaps_xchg_*): the folded partial sums
travel through peer-mapped mailboxes instead of an NCCL all-reduce.

* protocol on ONE device: two "ranks" = two streams + two mailboxes in one process (peer pointers are then plain
  pointers); every rank must produce scalars bit-identical to the other rank's and equal (fp64 re-association
  only) to the unsharded single-rank call, over several epochs (both phases);
* a missing peer must time out with NaN scalars and a latched status, not hang;
* two real processes over NCCL + CUDA IPC when the box has >= 2 GPUs (skipped otherwise; `bench.py --gpus 2`
  exercises the same path).
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

K, CC, B, H, W = 4, 4, 8, 64, 64
MIX = [0.15, 0.35, 0.2, 0.3]
CW1, CW2 = 0.07, 0.1


def _logits(dev, seed=5):
    g = torch.Generator(device=dev).manual_seed(seed)
    return [torch.randn(B, CC, H, W, generator=g, device=dev) * 2 for _ in range(K)]


def _single(L, lib, z, labels=None):
    dev = z[0].device
    ws = torch.zeros(lib.uaps_loss_workspace_bytes(K, CC), dtype=torch.uint8, device=dev)
    sums = torch.empty(lib.uaps_loss_sums_count(K, CC), dtype=torch.float64, device=dev)
    sc = torch.empty(lib.uaps_loss_scalars_count(K, CC), dtype=torch.float32, device=dev)
    L.check(lib.uaps_loss_pass1_scalars(L.ptr_array(z), K, z[0].shape[0], CC, H * W, L.float_array(MIX),
                                        None if labels is None else labels.data_ptr(), ws.data_ptr(), sums.data_ptr(),
                                        None, None, 0, CW1, CW2, sc.data_ptr(), None, L.stream_ptr()), "single")
    torch.cuda.synchronize()
    return sc, sums


@pytest.mark.parametrize("supervised", [False, True])
def test_two_ranks_on_one_device(supervised):
    from uaps_b200 import _lib as L
    lib = L.lib()
    dev = torch.device("cuda:0")
    z = _logits(dev)
    labels = torch.randint(0, CC, (B, H, W), device=dev) if supervised else None
    ref_sc, ref_sums = _single(L, lib, z, labels)
    world = 2
    boxes = (C.c_void_p * world)()
    for r in range(world):
        p = C.c_void_p()
        L.check(lib.uaps_xchg_alloc(C.byref(p)), "alloc")
        boxes[r] = p.value
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    shard = B // world
    ws = [torch.zeros(lib.uaps_loss_workspace_bytes(K, CC), dtype=torch.uint8, device=dev) for _ in range(world)]
    sums = [torch.empty(lib.uaps_loss_sums_count(K, CC), dtype=torch.float64, device=dev) for _ in range(world)]
    sc = [torch.empty(lib.uaps_loss_scalars_count(K, CC), dtype=torch.float32, device=dev) for _ in range(world)]
    zs = [[t[r * shard:(r + 1) * shard].contiguous() for t in z] for r in range(world)]
    ls = [None if labels is None else labels[r * shard:(r + 1) * shard].contiguous() for r in range(world)]
    torch.cuda.synchronize()
    try:
        for epoch in range(1, 6):                      # both phases, several times
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    L.check(lib.uaps_loss_pass1_exchange(L.ptr_array(zs[r]), K, shard, CC, H * W, L.float_array(MIX),
                                                         None if ls[r] is None else ls[r].data_ptr(), ws[r].data_ptr(),
                                                         sums[r].data_ptr(), None, None, 0, boxes, r, world, epoch, B * H * W,
                                                         CW1, CW2, sc[r].data_ptr(), None, None, streams[r].cuda_stream), "exchange")
            torch.cuda.synchronize()
            assert torch.equal(sc[0], sc[1]), "ranks must finalize bit-identical scalars"
            assert torch.equal(sums[0], sums[1])
            assert torch.isfinite(sc[0]).all()
            torch.testing.assert_close(sums[0], ref_sums, rtol=1e-12, atol=1e-9)       # fp64 re-association only
            torch.testing.assert_close(sc[0], ref_sc, rtol=1e-6, atol=1e-9)
            out = C.c_uint(7)
            L.check(lib.uaps_xchg_status(boxes[0], C.byref(out), L.stream_ptr()), "status")
            assert out.value == 0
    finally:
        torch.cuda.synchronize()
        for r in range(world):
            lib.uaps_xchg_free(boxes[r])


def test_missing_peer_times_out_instead_of_hanging(monkeypatch):
    from uaps_b200 import _lib as L
    lib = L.lib()
    dev = torch.device("cuda:0")
    z = _logits(dev)
    monkeypatch.setenv("UAPS_XCHG_TIMEOUT_MS", "30")
    boxes = (C.c_void_p * 2)()
    for r in range(2):
        p = C.c_void_p()
        L.check(lib.uaps_xchg_alloc(C.byref(p)), "alloc")
        boxes[r] = p.value
    ws = torch.zeros(lib.uaps_loss_workspace_bytes(K, CC), dtype=torch.uint8, device=dev)
    sums = torch.empty(lib.uaps_loss_sums_count(K, CC), dtype=torch.float64, device=dev)
    sc = torch.zeros(lib.uaps_loss_scalars_count(K, CC), dtype=torch.float32, device=dev)
    try:
        L.check(lib.uaps_loss_pass1_exchange(L.ptr_array(z), K, B, CC, H * W, L.float_array(MIX), None, ws.data_ptr(),
                                             sums.data_ptr(), None, None, 0, boxes, 0, 2, 1, 2 * B * H * W, CW1, CW2,
                                             sc.data_ptr(), None, None, L.stream_ptr()), "exchange")     # rank 1 never shows up
        torch.cuda.synchronize()
        assert torch.isnan(sc).all()
        out = C.c_uint(0)
        L.check(lib.uaps_xchg_status(boxes[0], C.byref(out), L.stream_ptr()), "status")
        assert out.value == 1
    finally:
        for r in range(2):
            lib.uaps_xchg_free(boxes[r])


def test_invalid_exchange_arguments():
    from uaps_b200 import _lib as L
    lib = L.lib()
    dev = torch.device("cuda:0")
    z = _logits(dev)
    ws = torch.zeros(lib.uaps_loss_workspace_bytes(K, CC), dtype=torch.uint8, device=dev)
    sums = torch.empty(lib.uaps_loss_sums_count(K, CC), dtype=torch.float64, device=dev)
    sc = torch.zeros(lib.uaps_loss_scalars_count(K, CC), dtype=torch.float32, device=dev)
    boxes = (C.c_void_p * 2)()
    args = lambda world, rank, epoch, b=boxes: (L.ptr_array(z), K, B, CC, H * W, L.float_array(MIX), None, ws.data_ptr(),
                                                sums.data_ptr(), None, None, 0, b, rank, world, epoch, B * H * W, CW1, CW2,
                                                sc.data_ptr(), None, None, L.stream_ptr())
    assert lib.uaps_loss_pass1_exchange(*args(2, 0, 1)) != 0          # null mailboxes
    assert lib.uaps_loss_pass1_exchange(*args(9, 0, 1)) != 0          # too many ranks
    assert lib.uaps_loss_pass1_exchange(*args(2, 2, 1)) != 0          # rank out of range
    assert lib.uaps_loss_pass1_exchange(*args(2, 0, 0)) != 0          # epoch 0 is the "never written" flag value
    assert lib.uaps_xchg_mailbox_bytes() % 128 == 0


# ---- two processes, two GPUs ---------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from uaps_b200 import comm
        from uaps_b200.losses import uaps_unlabeled_loss
        zfull = _logits(dev)                                   # same seed on every rank -> same full batch
        shard = B // world
        res = []
        for mode in ("peer", "nccl"):
            os.environ["UAPS_LOSS_EXCHANGE"] = mode
            comm.close_all()
            for it in range(3):
                z = [t[rank * shard:(rank + 1) * shard].clone().requires_grad_(True) for t in zfull]
                loss, ps, unc, pseudo, _ = uaps_unlabeled_loss(z, MIX, CW1, CW2, group=dist.group.WORLD, return_pseudo=True)
                loss.backward()
            used_peer = comm.exchange_for(dist.group.WORLD, dev) is not None
            res.append((mode, used_peer, loss.item(), ps.item(), unc.item(), z[0].grad.double().abs().sum().item()))
        comm.close_all()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_processes_peer_exchange_matches_nccl_and_single():
    import torch.multiprocessing as mp
    from uaps_b200.losses import uaps_unlabeled_loss
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    dev = torch.device("cuda:0")
    z = [t.requires_grad_(True) for t in _logits(dev)]
    loss, ps, unc, _, _ = uaps_unlabeled_loss(z, MIX, CW1, CW2)
    for rank in (0, 1):
        (m0, peer0, l0, p0, u0, _), (m1, peer1, l1, p1, u1, _) = out[rank]
        assert (m0, peer0, m1, peer1) == ("peer", True, "nccl", False)
        np.testing.assert_allclose([l0, p0, u0], [l1, p1, u1], rtol=1e-6)
        np.testing.assert_allclose([l0, p0, u0], [loss.item(), ps.item(), unc.item()], rtol=2e-6)
    assert out[0][0][2:5] == out[1][0][2:5]                  # both ranks: identical scalars


# ---- two processes: a sharded TRAINING ITERATION equals the reference's DataParallel semantics ------------------------
def _train_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle.unet_ref import feature_shapes, synthetic_rand, synthetic_state_dict
        from uaps_b200.train import UAPSConfig, UAPSTrainer
        from uaps_b200.unet import UNet_UAPS
        torch.backends.cudnn.allow_tf32 = False
        Bs, Hh, Cc = 2, 64, 4
        # (a) fp32 reference-precision path, every draw injected: must equal the single-process emulation
        model = UNet_UAPS(3, Cc, compute="fp32")
        model.load_state_dict(synthetic_state_dict(3, Cc, seed=11 + 7 * rank))     # ranks start DIFFERENT on purpose ...
        model = model.to(dev)
        tr = UAPSTrainer(model, UAPSConfig(num_classes=Cc), group=dist.group.WORLD)   # ... the constructor broadcasts rank 0's
        g = torch.Generator().manual_seed(2)
        xl, xu = torch.randn(world * Bs, 3, Hh, Hh, generator=g), torch.randn(world * Bs, 3, Hh, Hh, generator=g)
        yl = torch.randint(0, Cc, (world * Bs, Hh, Hh), generator=g)
        sl = slice(rank * Bs, (rank + 1) * Bs)
        to = lambda r: {k: [v.to(dev) if torch.is_tensor(v) else v for v in vals] for k, vals in r.items()}
        rl, ru = to(synthetic_rand(feature_shapes(Bs, Hh, Hh), 1 + rank)), to(synthetic_rand(feature_shapes(Bs, Hh, Hh), 5 + rank))
        out = tr.step(xl[sl].to(dev), yl[sl].to(dev), xu[sl].to(dev), mix_w=MIX, rand_l=rl, rand_u=ru)
        flat = tr.optimizer.flat_p.detach().cpu()
        # (b) bf16 captured path: replicas stay bit-identical and the loss falls
        torch.manual_seed(100 + rank)                                          # different initial weights per rank again
        m16 = UNet_UAPS(3, Cc).to(dev)
        t16 = UAPSTrainer(m16, UAPSConfig(num_classes=Cc, graph_warmup=1), group=dist.group.WORLD)
        yb = ((xl[sl, 0] > 0).long() + 2 * (xl[sl, 1] > 0).long()).to(dev)
        losses = [float(t16.step(xl[sl].to(dev), yb, xu[sl].to(dev))["loss"]) for _ in range(8)]
        q.put((rank, float(out["loss"]), flat, losses, t16.optimizer.flat_p.detach().cpu(), len(t16._graphs), t16.state is not None,
               t16.skipped_steps()))
    finally:
        from uaps_b200 import comm
        comm.close_all()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_training_iteration_matches_dataparallel_semantics():
    """Reference semantics (nn.DataParallel, UAPS_model.py:13): each replica runs the network on its shard with its OWN
    BatchNorm statistics, the logits are gathered and the losses (CE / Dice / mean exp(-KL)) are taken over the WHOLE batch,
    gradients are summed.  Emulated here in one process with the same injected draws and compared with two real ranks."""
    import torch.multiprocessing as mp
    from oracle.unet_ref import feature_shapes, synthetic_rand, synthetic_state_dict
    from uaps_b200.losses import uaps_supervised_loss, uaps_unlabeled_loss
    from uaps_b200.train import FlatAdam, FlatGradBuffer
    from uaps_b200.unet import UNet_UAPS
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        r = q.get(timeout=600)
        got[r[0]] = r[1:]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # single-process emulation on GPU 0
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    world, Bs, Hh, Cc = 2, 2, 64, 4
    model = UNet_UAPS(3, Cc, compute="fp32")
    model.load_state_dict(synthetic_state_dict(3, Cc, seed=11))            # rank 0's weights (what the broadcast distributes)
    model = model.to(dev).train()
    buf = FlatGradBuffer(model.parameters())
    opt = FlatAdam(buf, lr=1e-3)
    g = torch.Generator().manual_seed(2)
    xl, xu = torch.randn(world * Bs, 3, Hh, Hh, generator=g), torch.randn(world * Bs, 3, Hh, Hh, generator=g)
    yl = torch.randint(0, Cc, (world * Bs, Hh, Hh), generator=g)
    to = lambda r: {k: [v.to(dev) if torch.is_tensor(v) else v for v in vals] for k, vals in r.items()}
    outs_l, outs_u = [], []
    for r in range(world):
        sl = slice(r * Bs, (r + 1) * Bs)
        rl, ru = to(synthetic_rand(feature_shapes(Bs, Hh, Hh), 1 + r)), to(synthetic_rand(feature_shapes(Bs, Hh, Hh), 5 + r))
        outs_l.append(model(xl[sl].to(dev), rand=rl))
        outs_u.append(model(xu[sl].to(dev), rand=ru))
    cat = lambda outs: [torch.cat([o[k] for o in outs], 0) for k in range(4)]
    sup = uaps_supervised_loss(cat(outs_l), yl.to(dev))[0]
    from uaps_b200.ramps import get_current_consistency_weight
    cw = get_current_consistency_weight(0, 0.1, 200.0, 80)
    lu = uaps_unlabeled_loss(cat(outs_u), MIX, cw, cw)[0]
    loss = sup + lu
    buf.zero()
    loss.backward()
    opt.step()
    ref_flat = opt.flat_p.detach().cpu()
    for rank in (0, 1):
        loss_r, flat_r, losses16, flat16, ngraphs, dev_mode, skipped = got[rank]
        assert loss_r == pytest.approx(float(loss.detach()), rel=1e-5)
        # Adam normalises the gradient: compare the update direction of every parameter element that moved
        assert dev_mode and ngraphs == 1 and skipped == 0
        assert all(np.isfinite(losses16)) and losses16[-1] < losses16[0], losses16
    assert torch.equal(got[0][1], got[1][1]), "fp32 path: replicas diverged"
    assert torch.equal(got[0][3], got[1][3]), "bf16 captured path: replicas diverged"
    assert got[0][2] == got[1][2], "both ranks must log the same (global) loss"
    # parameters after one step vs the emulation: same update sign on > 99.5 % of the elements that moved, and close in value
    torch.testing.assert_close(got[0][1], ref_flat, rtol=0, atol=2.5e-3)           # |update| <= lr = 1e-3 per element
    m0 = UNet_UAPS(3, Cc, compute="fp32"); m0.load_state_dict(synthetic_state_dict(3, Cc, seed=11)); m0 = m0.to(dev)
    start = FlatAdam(FlatGradBuffer(m0.parameters()), lr=1e-3).flat_p.detach().cpu()
    du_ref, du_got = ref_flat - start, got[0][1] - start
    live = du_ref.abs() > 5e-4                                                  # elements with a clear (non-noise) gradient
    agree = (torch.sign(du_ref[live]) == torch.sign(du_got[live])).float().mean().item()
    assert agree > 0.995, agree
