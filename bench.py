#!/usr/bin/env python
"""bench.py -- the UAPS unlabeled-batch hot path on B200, one JSON line (contract: task brief (4)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-train-step]

Headline workload (BASELINE.json configs[1], the shape of configs[2]'s unlabeled batch): the fused
pseudo-label + KL-uncertainty + weighted CE/Dice loss, forward + backward, K=4 decoders, C=4
classes, 64 x 256x256 pixels per GPU (268 MB of fp32 logits, > the 126 MB L2).  One "step" = pass 1 +
fold/finalize + pass 2 over one batch (3 launches chained by programmatic dependent launch).  metric =
pixels/s over all ranks (weak scaling: every rank has its own batch; the <=70-double partial-sum vector is
exchanged between the passes inside the fold kernel, through NVLink peer-memory mailboxes).

  value     : logits resident in HBM, C-ABI calls timed with CUDA events on the launching stream.
  e2e       : the public API (uaps_unlabeled_loss + backward) fed from pinned HOST buffers, H2D copy
              of the logits and D2H read of the loss scalars inside the timed region.
  roofline  : the dominant kernel (pass 2: 8*K*C algorithmic bytes / pixel) against the measured HBM
              copy bandwidth in MEASURED_PEAKS.json; per-kernel numbers under "kernels".
  cpu_baseline : the oracle (restated reference expressions, torch CPU) on the box's host cores.
  eager_cuda_baseline : the same expressions as torch-eager ops on the same GPU (the incumbent GPU path).
  train_step: secondary -- full UAPS iteration (two forwards, both losses, backward, Adam) iters/s.
  inference : validation forward (BatchNorm folded, LeakyReLU in the conv epilogue), throughput and batch-1 latency.

--impl reference times the reference's CPU implementation of the same path (the oracle port; the
reference itself is pure Python over torch and its loss section is inline code that cannot be
imported -- see DESIGN.md) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, C, B, H, W = 4, 4, 64, 256, 256
CW1 = CW2 = 0.1
WORKLOAD = f"fused_loss_fwd_bwd K={K} C={C} B={B}/gpu {H}x{W} fp32-logits"
CPU_SAMPLE_B = 8                       # bounded CPU sample: 8 x 256 x 256 pixels per step
TRAIN_B = 64                           # labeled + unlabeled images per GPU (BASELINE configs[2]: batch 64+64)
TRAIN_B_FP32 = 16                      # the fp32 cuDNN comparison path runs a quarter batch (it is ~3x slower)


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes per pass-2 launch from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "loss_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if mx and s > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def eager_cuda_baseline(dev, steps: int = 5, warmup: int = 2):
    """BASELINE.md section 2, "also reported": the reference's own loss expressions (the oracle restatement, bit-identical
    to the executed reference lines) as torch-eager ops ON THE SAME B200, same workload as the headline, fwd + bwd.
    The incumbent GPU number the fused kernels replace -- a baseline leg, never part of the product path."""
    from oracle.uaps_loss_ref import unlabeled_loss_ref
    gen = torch.Generator(device=dev).manual_seed(1337)
    z = [(torch.randn(B, C, H, W, generator=gen, device=dev) * 2).requires_grad_(True) for _ in range(K)]
    mix_w = np.random.default_rng(1337).dirichlet(np.ones(K))

    def one():
        for t in z:
            t.grad = None
        unlabeled_loss_ref(z, mix_w, CW1, CW2)["loss_u"].backward()

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del z
    torch.cuda.empty_cache()
    return {"value": B * H * W / (ms * 1e-3), "unit": "pixels/s", "ms_per_step": ms,
            "what": "reference expressions (UAPS_train.py:186-189, 223-282) as torch-eager CUDA ops on this GPU, fwd+bwd, same workload"}


def cpu_baseline(steps: int, warmup: int, threads: int):
    """The oracle's unlabeled loss fwd+bwd on CPU torch (the reference's expressions), pixels/s."""
    from oracle.uaps_loss_ref import unlabeled_loss_ref
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1337)
    z = [(torch.randn(CPU_SAMPLE_B, C, H, W, generator=g) * 2).requires_grad_(True) for _ in range(K)]
    mix_w = np.random.default_rng(1337).dirichlet(np.ones(K))
    times = []
    for i in range(warmup + steps):
        for t in z:
            t.grad = None
        t0 = time.perf_counter()
        out = unlabeled_loss_ref(z, mix_w, CW1, CW2)
        out["loss_u"].backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    px = CPU_SAMPLE_B * H * W
    return px * len(times) / sum(times), sum(times) / len(times)


def cpu_train_step(steps: int, warmup: int, threads: int, batch: int = 4):
    """The reference's full iteration on CPU (configs[0]): oracle functional U-Net + oracle losses + Adam."""
    from oracle.unet_ref import feature_shapes, synthetic_rand, synthetic_state_dict, unet_uaps_ref
    from oracle.uaps_loss_ref import supervised_loss_ref, unlabeled_loss_ref
    torch.set_num_threads(threads)
    sd = synthetic_state_dict(3, C)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    full = {**sd, **params}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    g = torch.Generator().manual_seed(1337)
    xl, xu = torch.randn(batch, 3, H, W, generator=g), torch.randn(batch, 3, H, W, generator=g)
    yl = torch.randint(0, C, (batch, H, W), generator=g)
    rand = synthetic_rand(feature_shapes(batch, H, W))
    mix_w = np.random.default_rng(1337).dirichlet(np.ones(K))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ol = unet_uaps_ref(xl, full, rand)[:4]
        ou = unet_uaps_ref(xu, full, rand)[:4]
        loss = supervised_loss_ref(ol, yl)["supervised_loss"] + unlabeled_loss_ref(ou, mix_w, CW1, CW2)["loss_u"]
        opt.zero_grad()
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return len(times) / sum(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    pxs, sec = cpu_baseline(args.steps, min(args.warmup, 2), threads)
    line = {
        "impl": "reference", "metric": "fused_pl_kl_loss_pixels_per_s", "value": pxs, "unit": "pixels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "K": K, "C": C, "H": H, "W": W},
        "cpu_baseline": {"value": pxs, "unit": "pixels/s", "cores": threads, "kind": "port",
                         "sample": f"oracle (reference expressions, torch CPU) fwd+bwd on {CPU_SAMPLE_B}x{H}x{W} px per step"},
        "e2e": {"value": pxs, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_train_step:
        try:
            line["train_step"] = {"iters_per_s": cpu_train_step(2, 1, threads), "unit": "iters/s",
                                  "config": f"UNet_UAPS 3x{H}x{W} C={C} K={K} batch 4+4 (BASELINE configs[0]), torch CPU"}
        except Exception as e:                                   # the headline line must still print
            line["train_step"] = {"error": repr(e)[:200]}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from uaps_b200 import _lib as L
    from uaps_b200.losses import uaps_unlabeled_loss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: uaps_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    lib = L.lib()
    N = B * H * W
    gen = torch.Generator(device=dev).manual_seed(1337 + rank)
    z = [torch.randn(B, C, H, W, generator=gen, device=dev) * 2 for _ in range(K)]
    dz = [torch.empty_like(t) for t in z]
    mix_w = np.random.default_rng(1337).dirichlet(np.ones(K))
    ws = torch.zeros(lib.uaps_loss_workspace_bytes(K, C), dtype=torch.uint8, device=dev)
    sums = torch.empty(lib.uaps_loss_sums_count(K, C), dtype=torch.float64, device=dev)
    sc = torch.empty(lib.uaps_loss_scalars_count(K, C), dtype=torch.float32, device=dev)
    go = torch.zeros_like(sc)
    go[0] = 1.0
    zp, dzp, w_arr = L.ptr_array(z), L.ptr_array(dz), L.float_array(mix_w)
    st = L.stream_ptr()

    xchg = None
    if world > 1:                # the loss sums travel through NVLink peer mailboxes (NCCL only if that is unavailable)
        from uaps_b200.comm import exchange_for
        xchg = exchange_for(group, dev)

    def step(ev=None):
        # ev = (after pass 1 + fold/finalize [+ all-reduce], after pass 2); the step starts where the previous one ended
        if world == 1:          # single rank: fold + finalize fused into one launch behind pass 1
            L.check(lib.uaps_loss_pass1_scalars(zp, K, B, C, H * W, w_arr, None, ws.data_ptr(), sums.data_ptr(), None, None, 0,
                                                CW1, CW2, sc.data_ptr(), None, st), "pass1")
        elif xchg is not None:  # fold + peer-memory exchange + finalize in one launch
            L.check(lib.uaps_loss_pass1_exchange(zp, K, B, C, H * W, w_arr, None, ws.data_ptr(), sums.data_ptr(), None, None, 0,
                                                 xchg.ptrs, rank, world, xchg.next_epoch(), N * world, CW1, CW2, sc.data_ptr(), None, None, st),
                    "pass1")
        else:
            L.check(lib.uaps_loss_pass1(zp, K, B, C, H * W, w_arr, None, ws.data_ptr(), sums.data_ptr(), None, None, 0, st), "pass1")
            dist.all_reduce(sums, group=group)
            L.check(lib.uaps_loss_finalize(sums.data_ptr(), K, C, N * world, CW1, CW2, 0, sc.data_ptr(), st), "finalize")
        if ev: ev[0].record()
        L.check(lib.uaps_loss_pass2(zp, K, B, C, H * W, w_arr, None, sc.data_ptr(), go.data_ptr(), dzp, 0, None, st), "pass2")
        if ev: ev[1].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: device-resident inputs ---------------------------------------------------
    # Per-kernel events on every EV_EVERY-th step only (three records: start, after pass 1 + fold/finalize, end): an event
    # record between two launches cancels their programmatic-dependent-launch overlap, so the other steps run as the
    # library is used in a training loop.  `value` is from the outer pair of events around all K steps.
    EV_EVERY = 4
    evs = {i: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for i in range(0, args.steps, EV_EVERY)}
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        ev = evs.get(i)
        if ev:
            ev[0].record()
        step(ev[1:] if ev else None)
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    t1 = sum(e[0].elapsed_time(e[1]) for e in evs.values()) / len(evs)      # pass 1 + fold/finalize (+ exchange)
    t2 = sum(e[1].elapsed_time(e[2]) for e in evs.values()) / len(evs)      # pass 2

    # ---- e2e: public API, pinned host logits -> H2D -> fwd + bwd -> D2H loss scalars ---------------
    zh = [torch.empty((B, C, H, W), dtype=torch.float32).pin_memory() for _ in range(K)]
    for a, b in zip(zh, z):
        a.copy_(b)
    zd = [torch.empty_like(t).requires_grad_(True) for t in z]
    out_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_step():
        for hsrc, d in zip(zh, zd):
            d.grad = None
            d.data.copy_(hsrc, non_blocking=True)
        loss, ps, unc, _, _ = uaps_unlabeled_loss(zd, mix_w, CW1, CW2, group=group)
        loss.backward()
        out_host.copy_(torch.stack([loss.detach(), ps.detach(), unc.detach()]), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks ------------------------------------------------------------------------
    t = torch.tensor([ms_total, ms_e2e, t1, t2], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, t1, t2 = t.tolist()

    # secondary shape (DAGM-sized images): same kernels, 4x the pixels per launch
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = loss_sweep(dev, lib, L)

    train = None
    if not args.no_train_step:
        try:
            train = train_step_bench(dev, group, world, rank, compute="bf16")
            train["fp32_path"] = train_step_bench(dev, group, world, rank, compute="fp32", batch=TRAIN_B_FP32)
        except Exception as e:
            train = {"error": repr(e)[:300]}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        ms_step = ms_total / args.steps
        value = N * world * args.steps / (ms_total * 1e-3)
        bytes1, bytes2 = 4 * K * C * N, 8 * K * C * N
        ach2 = bytes2 / (t2 * 1e-3) / 1e9
        ach1 = bytes1 / (t1 * 1e-3) / 1e9
        traffic = ncu_traffic()
        line = {
            "metric": "fused_pl_kl_loss_pixels_per_s", "value": value, "unit": "pixels/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "K": K, "C": C, "H": H, "W": W, "batch_per_gpu": B,
                       "l2": "inputs (268 MB logits + 268 MB gradients per GPU) larger than the 126 MB L2; no flush",
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "exchange": None if world == 1 else ("nvlink peer-memory mailboxes, fused into the fold kernel"
                                                            if xchg is not None else "nccl all-reduce of the sums")},
            "roofline": {"bound": "hbm", "kernel": "loss_pass2_kernel<K=4,C=4,VEC=2,PF> (uaps_loss_pass2)", "achieved": ach2, "peak": peak,
                         "unit": "GB/s", "frac": ach2 / peak, "peak_source": peak_src,
                         "traffic": None if not traffic else traffic.get("pass2_dram_bytes_per_launch"),
                         "algorithmic_bytes_per_launch": bytes2,
                         "fwd_bwd_frac": (bytes1 + bytes2) / (ms_step * 1e-3) / 1e9 / peak},
            "kernels": {"pass1_fold_finalize" + ("_exchange" if world > 1 else ""):
                            {"ms": t1, "GBps": ach1, "frac": ach1 / peak, "algorithmic_bytes": bytes1},
                        "pass2": {"ms": t2, "GBps": ach2, "frac": ach2 / peak, "algorithmic_bytes": bytes2}},
            "e2e": {"value": N * world * e2e_steps / (ms_e2e * 1e-3), "unit": "pixels/s",
                    "h2d_bytes_per_step": 4 * K * C * N, "d2h_bytes_per_step": 12, "steps": e2e_steps,
                    "api": "uaps_b200.losses.uaps_unlabeled_loss + backward, pinned host logits"},
            "kernel_events": f"sampled on every {EV_EVERY}th step ({len(evs)} of {args.steps})",
            "gpu_launches": (3 if (world == 1 or xchg is not None) else 4) * args.steps,   # pass1, fold(+exchange)+finalize, pass2
            "clocks": clocks,
        }
        if world == 1:
            threads = os.cpu_count() or 1
            pxs, sec = cpu_baseline(5, 1, threads)
            line["cpu_baseline"] = {"value": pxs, "unit": "pixels/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle (reference expressions, torch CPU) fwd+bwd, 5 steps of {CPU_SAMPLE_B}x{H}x{W} px"}
            try:
                line["eager_cuda_baseline"] = eager_cuda_baseline(dev)
            except Exception as e:                       # noqa: BLE001 -- a baseline leg must not sink the bench line
                line["eager_cuda_baseline"] = {"error": repr(e)[:200]}
        if sweep is not None:
            line["sweep"] = sweep
        if train is not None:
            line["train_step"] = train
        if world == 1 and not args.no_train_step:
            try:
                line["inference"] = inference_bench(dev)
            except Exception as e:                       # noqa: BLE001
                line["inference"] = {"error": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def loss_sweep(dev, lib, L, iters: int = 10):
    """Other points of BASELINE configs[1]'s sweep, device-resident, same timing method as the headline."""
    peak, _ = measured_hbm_peak()
    out = []
    for (k, c, b, h, w) in [(4, 4, 64, 512, 512), (4, 2, 32, 512, 512), (5, 2, 32, 256, 512), (6, 4, 32, 512, 512),
                            (4, 4, 8, 200, 200)]:
        z = [torch.randn(b, c, h, w, device=dev) * 2 for _ in range(k)]
        dz = [torch.empty_like(t) for t in z]
        ws = torch.zeros(lib.uaps_loss_workspace_bytes(k, c), dtype=torch.uint8, device=dev)
        sums = torch.empty(lib.uaps_loss_sums_count(k, c), dtype=torch.float64, device=dev)
        sc = torch.empty(lib.uaps_loss_scalars_count(k, c), dtype=torch.float32, device=dev)
        go = torch.zeros_like(sc); go[0] = 1.0
        zp, dzp, wa, st = L.ptr_array(z), L.ptr_array(dz), L.float_array([1.0 / k] * k), L.stream_ptr()
        n = b * h * w

        def step():
            L.check(lib.uaps_loss_pass1_scalars(zp, k, b, c, h * w, wa, None, ws.data_ptr(), sums.data_ptr(), None, None, 0,
                                                CW1, CW2, sc.data_ptr(), None, st), "p1")
            L.check(lib.uaps_loss_pass2(zp, k, b, c, h * w, wa, None, sc.data_ptr(), go.data_ptr(), dzp, 0, None, st), "p2")
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out.append({"K": k, "C": c, "B": b, "H": h, "W": w, "ms_per_step": ms, "pixels_per_s": n / (ms * 1e-3),
                    "fwd_bwd_frac_of_hbm_peak": 12 * k * c * n / (ms * 1e-3) / 1e9 / peak})
        del z, dz
    return out


def train_step_bench(dev, group, world, rank, steps: int = 5, warmup: int = 3, compute: str = "bf16", batch: int = TRAIN_B):
    """Secondary: full UAPS iteration at the NEU shape (3x256x256, C=4, K=4), batch + batch images per GPU."""
    import torch.distributed as dist
    from uaps_b200.train import UAPSTrainer
    from uaps_b200.unet import UNet_UAPS
    torch.manual_seed(1337)
    model = UNet_UAPS(3, C, compute=compute).to(dev)
    trainer = UAPSTrainer(model, group=group)
    gen = torch.Generator().manual_seed(1337 + rank)
    xl_h = torch.randn(batch, 3, H, W, generator=gen).pin_memory()
    xu_h = torch.randn(batch, 3, H, W, generator=gen).pin_memory()
    yl_h = torch.randint(0, C, (batch, H, W), generator=gen).pin_memory()
    loss_h = torch.empty((), dtype=torch.float32).pin_memory()

    def it():
        xl, xu, yl = xl_h.to(dev, non_blocking=True), xu_h.to(dev, non_blocking=True), yl_h.to(dev, non_blocking=True)
        out = trainer.step(xl, yl, xu)
        loss_h.copy_(out["loss"], non_blocking=True)

    for _ in range(warmup):
        it()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        it()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    return {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "images_per_s": 2 * batch * world * 1e3 / ms, "unit": "iters/s",
            "loss": float(loss_h), "config": f"UNet_UAPS 3x{H}x{W} C={C} K={K}, {batch}+{batch} images/GPU, "
            f"dp{world}, host images in the timed region (e2e)",
            "kernels": "bf16 channels-last: tcgen05 implicit-GEMM fprop/dgrad/wgrad, fused BN+LeakyReLU+dropout, pool/upsample, Philox perturbations, fused losses (all hand-written sm_100a); torch: autograd glue, fused Adam"
            if compute == "bf16" else "cuDNN fp32 (reference-precision path)"}


def inference_bench(dev, batch: int = 64, steps: int = 10, warmup: int = 3):
    """Row f3: validation / inference forward (UAPS_train.py:367-393) -- main decoder only, BatchNorm folded into the conv
    weights, LeakyReLU in the conv epilogue.  The reference's README quotes 4.48 ms / 256x256 image for the main decoder
    (fig_data/decoder-effect.jpg, hardware not stated); reported beside, not as vs_baseline (different metric)."""
    from uaps_b200.unet import UNet_UAPS
    torch.manual_seed(1337)
    model = UNet_UAPS(3, C, compute="bf16").to(dev).eval()
    x_h = torch.randn(batch, 3, H, W).pin_memory()
    lab_h = torch.empty((batch, H, W), dtype=torch.uint8).pin_memory()

    def it():
        logits = model.predict(x_h.to(dev, non_blocking=True))
        lab_h.copy_(logits.argmax(1).to(torch.uint8), non_blocking=True)      # the label map goes back to the host

    for _ in range(warmup):
        it()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        it()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps

    # latency mode: one image per call, host image in, label map out; plain launches vs the captured CUDA graph
    x1_h = torch.randn(1, 3, H, W).pin_memory()
    lab1_h = torch.empty((1, H, W), dtype=torch.uint8).pin_memory()
    lat = {}
    for name, fn in (("launches", model.predict), ("cuda_graph", model.predict_graphed)):
        def one():
            lab1_h.copy_(fn(x1_h.to(dev, non_blocking=True)).argmax(1).to(torch.uint8), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(5):
            one()
        t0 = time.perf_counter()
        for _ in range(50):
            one()
        lat[name] = (time.perf_counter() - t0) / 50 * 1e3
    return {"ms_per_image": ms / batch, "images_per_s": batch * 1e3 / ms, "batch": batch, "ms_per_batch": ms,
            "latency_ms_batch1": lat,
            "config": f"UNet_UAPS.predict 3x{H}x{W} C={C}, bf16 path, eval mode, host images in / uint8 label map out",
            "published_ms_per_image": 4.48, "published_source": "reference README Fig. 9 (hardware not stated)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-train-step", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
