"""World-size-2 gloo tests (CPU) of the N > 1 path's host logic: the partial-sum exchange of the fused
loss (layout of uaps_loss_sums_count, all-reduce, finalize formula) and the flat gradient all-reduce.
The CUDA kernels themselves are covered by the -m gpu tests; here each rank forms its shard's partial
sums with plain torch, the product's `_allreduce_sums` exchanges them, and the result must equal the
oracle evaluated on the whole batch -- the semantics the reference gets from DataParallel's gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.uaps_loss_ref import unlabeled_loss_ref

K, C, B, H, W = 4, 4, 4, 16, 16
MIX = [0.1, 0.2, 0.3, 0.4]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _logits():
    g = torch.Generator().manual_seed(77)
    return [torch.randn(B, C, H, W, generator=g) * 2 for _ in range(K)]


def _shard_sums(z):
    """The vector pass 1 produces for one shard: [K] sum(-logp[y]) | [K] sum E | [K] sum V | [KC] I | [KC] P | [C] T."""
    out = unlabeled_loss_ref(z, MIX, 0.1, 0.1)
    y = out["pseudo"]
    oh = torch.nn.functional.one_hot(y, C).permute(0, 3, 1, 2).double()
    ce = [-(torch.log_softmax(t.double(), 1) * oh).sum() for t in z]
    sE = [e.double().sum() for e in out["exp_var"]]
    sV = [v.double().sum() for v in out["var"]]
    I = [(p.double() * oh).sum((0, 2, 3)) for p in out["soft"]]
    P = [p.double().sum((0, 2, 3)) for p in out["soft"]]
    T = oh.sum((0, 2, 3))
    return torch.cat([torch.stack(ce), torch.stack(sE), torch.stack(sV), torch.cat(I), torch.cat(P), T])


def _finalize(sums, n, cw1, cw2):
    """Host mirror of loss_finalize_kernel (uaps_b200/csrc/fused_loss_impl.cuh)."""
    s = sums.double()
    sCE, sE, sV = s[:K], s[K:2 * K], s[2 * K:3 * K]
    sI, sP = s[3 * K:3 * K + K * C].view(K, C), s[3 * K + K * C:3 * K + 2 * K * C].view(K, C)
    sT = s[3 * K + 2 * K * C:]
    dice = 1 - (2 * sI / (sP + sT + 1e-7)).mean(1)
    ps = 0.5 * (sCE / n + dice)
    ps_loss = (ps * sE / n).mean()
    unc = (sV / n).mean()
    return cw1 * ps_loss + cw2 * unc, ps_loss, unc


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from uaps_b200.losses import _allreduce_sums
        from uaps_b200.train import FlatGradBuffer
        z = _logits()
        per = B // world
        shard = [t[rank * per:(rank + 1) * per] for t in z]
        sums = _shard_sums(shard)
        assert sums.numel() == 3 * K + 2 * K * C + C
        w = _allreduce_sums(sums, dist.group.WORLD)
        assert w == world
        loss, ps, unc = _finalize(sums, per * H * W * w, 0.1, 0.07)
        # flat gradient buffer: SUM all-reduce over ranks, views stay attached to the parameters
        lin = torch.nn.Linear(3, 2)
        buf = FlatGradBuffer(lin.parameters())
        for p in lin.parameters():
            p.grad.fill_(float(rank + 1))
        buf.all_reduce_sum(dist.group.WORLD)
        ok_grad = all(torch.all(p.grad == sum(range(1, world + 1))).item() for p in lin.parameters())
        ok_view = lin.weight.grad.data_ptr() == buf.flat.data_ptr()
        # a gloo group can never use the NVLink peer mailboxes: the loss must pick the all-reduce route on every rank
        from uaps_b200.comm import exchange_for
        ok_view = ok_view and exchange_for(dist.group.WORLD, torch.device("cuda", 0)) is None
        ret[rank] = (loss.item(), ps.item(), unc.item(), ok_grad, ok_view)
    finally:
        dist.destroy_process_group()


def test_sharded_partial_sums_reproduce_whole_batch_loss():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    ref = unlabeled_loss_ref(_logits(), MIX, 0.1, 0.07)          # whole batch on one "device"
    for rank in range(world):
        loss, ps, unc, ok_grad, ok_view = ret[rank]
        assert loss == pytest.approx(ref["loss_u"].item(), rel=1e-6)
        assert ps == pytest.approx(ref["ps_loss"].item(), rel=1e-6)
        assert unc == pytest.approx(ref["l_uncert"].item(), rel=1e-6)
        assert ok_grad and ok_view


def test_single_process_allreduce_is_identity():
    from uaps_b200.losses import _allreduce_sums
    s = torch.arange(5, dtype=torch.float64)
    assert _allreduce_sums(s, None) == 1 and torch.equal(s, torch.arange(5, dtype=torch.float64))


def test_flat_grad_buffer_layout():
    """Every parameter's gradient slot starts on a 256-byte boundary (the kernels take 16-byte vector loads of biases
    and BatchNorm parameters), the views alias the flat buffer, and padding stays zero under zero()."""
    from uaps_b200.train import FlatGradBuffer
    params = [torch.nn.Parameter(torch.randn(*s)) for s in [(16, 3, 3, 3), (16,), (2,), (4, 16, 3, 3), (7,)]]
    buf = FlatGradBuffer(params)
    assert buf.offsets == sorted(buf.offsets) and all(o % FlatGradBuffer.ALIGN == 0 for o in buf.offsets)
    for p, off in zip(params, buf.offsets):
        assert p.grad.shape == p.shape and p.grad.data_ptr() == buf.flat.data_ptr() + 4 * off
        p.grad.fill_(1.0)
    assert buf.flat.sum().item() == sum(p.numel() for p in params)      # padding untouched
    buf.zero()
    assert buf.flat.abs().sum().item() == 0.0
