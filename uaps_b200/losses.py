"""Loss surface of the UAPS hot path, backed by the fused sm_100a kernels.

Function forms of what the reference has inline in its training loop:

* ``uaps_unlabeled_loss``  <- UAPS_train.py:186-189, 223-282 (K softmaxes, mean prediction, KL
  uncertainty maps, Dirichlet-mixed argmax pseudo-label, uncertainty-weighted CE + Dice, ramped total);
* ``uaps_supervised_loss`` <- UAPS_train.py:194-218 (labeled batch, mean over decoders of 0.5(CE+Dice));
* ``dice_loss`` / ``ce_loss`` <- utilities/pytorch_losses.py:54-89 and ``CrossEntropyLoss()`` (:75),
  same argument order and shapes as the reference.

Every function is a ``torch.autograd.Function`` over the C ABI in include/uaps_b200.h: pass 1 and
the scalar finalize run in forward, pass 2 in backward.  Only the K logits tensors and a <=
(8 + 4K + 2KC)-float scalar vector are saved for backward.  CUDA tensors only; no fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L

_workspaces = {}
_WORKSPACE_CACHE_MAX = 32


def _workspace(device: torch.device, K: int, C: int) -> torch.Tensor:
    """Pass-1 scratch, zero-filled once per (device, stream, K, C) as the ABI requires.  The cache is bounded: the
    least recently used entry goes first (a long-lived process that keeps creating streams must not leak)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, K, C)
    ws = _workspaces.pop(key, None)
    if ws is None:
        nbytes = L.lib().uaps_loss_workspace_bytes(K, C)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        while len(_workspaces) >= _WORKSPACE_CACHE_MAX:
            _workspaces.pop(next(iter(_workspaces)))
    _workspaces[key] = ws                     # (re-)inserted last = most recently used
    return ws


def _prep_logits(logits: Sequence[torch.Tensor]) -> Tuple[List[torch.Tensor], int, int, int, int]:
    if len(logits) < 1 or len(logits) > L.KMAX:
        raise RuntimeError(f"number of decoders must be in [1, {L.KMAX}], got {len(logits)}")
    zs = []
    shape = logits[0].shape
    for z in logits:
        L.require_cuda(z)
        if z.dim() != 4 or z.shape != shape:
            raise RuntimeError("all logits must be [B, C, H, W] with the same shape")
        if z.dtype != torch.float32:
            raise RuntimeError("logits must be fp32 (the reference's decoders emit fp32)")
        zs.append(z if z.is_contiguous() else z.contiguous())
    B, C, H, W = shape
    if C < 2 or C > L.CMAX:
        raise RuntimeError(f"class count must be in [2, {L.CMAX}], got {C}")
    return zs, B, C, H, W


def _allreduce_sums(sums: torch.Tensor, group) -> int:
    """The path's one exchange step: sum the per-rank partial-sum vector (<= 3K+2KC+C doubles)."""
    import torch.distributed as dist
    if group is None or not dist.is_initialized():
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return world


class _FusedLossFn(torch.autograd.Function):
    """Shared autograd node for the unlabeled (labels=None) and supervised (labels given) modes."""

    @staticmethod
    def forward(ctx, mix_w, cw1, cw2, labels, group, want_pseudo, want_exp_var, flags, n_global, dstate, *logits):
        zs, B, C, H, W = _prep_logits(logits)
        K, HW = len(zs), H * W
        dev = zs[0].device
        lib = L.lib()
        sup = labels is not None
        if sup:
            L.require_cuda(labels)
            if labels.dtype != torch.int64 or tuple(labels.shape) != (B, H, W):
                raise RuntimeError("labels must be int64 [B, H, W]")
            labels = labels.contiguous()
            w_arr = None
        elif dstate is not None:                    # device-resident iteration: mix_w / cw1 / cw2 live in UapsStepState
            w_arr = None
        else:
            if mix_w is None or len(mix_w) != K:
                raise RuntimeError("mix_w must hold one weight per decoder")
            w_arr = L.float_array(mix_w)            # fp32 rounding = torch's python-scalar * tensor rule
        # dstate = (DeviceStepState, LossExchange | None, epoch offset inside the iteration)
        wcw_dev = dstate[0].ptr("mix_w") if (dstate is not None and not sup) else None
        with L.on_device(dev):
            sums = torch.empty(lib.uaps_loss_sums_count(K, C), dtype=torch.float64, device=dev)
            scalars = torch.empty(lib.uaps_loss_scalars_count(K, C), dtype=torch.float32, device=dev)
            pseudo = torch.empty((B, H, W), dtype=torch.int64, device=dev) if (want_pseudo and not sup) else None
            exp_var = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(K)] \
                if (want_exp_var and not sup) else None
            zp = L.ptr_array(zs)
            import torch.distributed as dist
            single = (group is None or not dist.is_initialized() or dist.get_world_size(group) == 1) and not n_global
            common = (zp, K, B, C, HW, w_arr, None if not sup else labels.data_ptr(),
                      _workspace(dev, K, C).data_ptr(), sums.data_ptr(),
                      None if pseudo is None else pseudo.data_ptr(),
                      None if exp_var is None else L.ptr_array(exp_var), flags)
            if dstate is not None and dstate[1] is not None:  # this trainer's own mailboxes, epochs counted on the device
                xchg = dstate[1]
                n_tot = int(n_global) if n_global else B * HW * xchg.world
                L.check(lib.uaps_loss_pass1_exchange(*common, xchg.ptrs, xchg.rank, xchg.world, int(dstate[2]), n_tot,
                                                     float(cw1), float(cw2), scalars.data_ptr(), wcw_dev,
                                                     dstate[0].ptr("xchg_base"), L.stream_ptr()), "uaps_loss_pass1_exchange")
            elif single:                             # no exchange between the passes: fold + finalize fused
                L.check(lib.uaps_loss_pass1_scalars(*common, float(cw1), float(cw2), scalars.data_ptr(), wcw_dev,
                                                    L.stream_ptr()), "uaps_loss_pass1_scalars")
            else:
                if dstate is not None:
                    raise RuntimeError("a device-resident iteration over several ranks needs the peer-memory exchange")
                from .comm import exchange_for
                xchg = exchange_for(group, dev) if dist.is_initialized() else None
                if xchg is not None:                 # fold + NVLink peer-memory exchange + finalize in one kernel
                    n_tot = int(n_global) if n_global else B * HW * xchg.world
                    L.check(lib.uaps_loss_pass1_exchange(*common, xchg.ptrs, xchg.rank, xchg.world, xchg.next_epoch(), n_tot,
                                                         float(cw1), float(cw2), scalars.data_ptr(), None, None,
                                                         L.stream_ptr()), "uaps_loss_pass1_exchange")
                else:
                    L.check(lib.uaps_loss_pass1(*common, L.stream_ptr()), "uaps_loss_pass1")
                    world = _allreduce_sums(sums, group)
                    n_tot = int(n_global) if n_global else B * HW * world
                    L.check(lib.uaps_loss_finalize(sums.data_ptr(), K, C, n_tot, float(cw1), float(cw2), int(sup),
                                                   scalars.data_ptr(), L.stream_ptr()), "uaps_loss_finalize")
        ctx.save_for_backward(scalars, *([labels] if sup else []), *zs)
        ctx.meta = (K, B, C, HW, sup, flags, tuple(float(w) for w in mix_w) if (not sup and dstate is None) else None, wcw_dev)
        extra = []
        if pseudo is not None:
            extra.append(pseudo)
        if exp_var is not None:
            extra.extend(exp_var)
        if extra:
            ctx.mark_non_differentiable(*extra)
        return (scalars, *extra)

    @staticmethod
    def backward(ctx, g_scalars, *_unused):
        K, B, C, HW, sup, flags, mix_w, wcw_dev = ctx.meta
        saved = ctx.saved_tensors
        scalars = saved[0]
        labels = saved[1] if sup else None
        zs = saved[2:] if sup else saved[1:]
        dev = scalars.device
        # the autograd gradient of the scalars vector IS the kernel's grad_out (same layout)
        grad_out = g_scalars.to(torch.float32).contiguous()
        with L.on_device(dev):
            dz = [torch.empty_like(z) for z in zs]
            L.check(L.lib().uaps_loss_pass2(L.ptr_array(zs), K, B, C, HW,
                                            None if (sup or mix_w is None) else L.float_array(mix_w),
                                            labels.data_ptr() if sup else None,
                                            scalars.data_ptr(), grad_out.data_ptr(), L.ptr_array(dz),
                                            flags, wcw_dev, L.stream_ptr()), "uaps_loss_pass2")
        return (None,) * 10 + tuple(dz)


def _split_extra(out, want_pseudo, want_exp_var):
    rest = list(out[1:])
    pseudo = rest.pop(0) if want_pseudo else None
    return pseudo, (rest if want_exp_var else None)


def _dev_ctx(step_state):
    """(DeviceStepState, LossExchange | None, epoch offset) for a loss call inside a device-resident iteration."""
    if step_state is None:
        return None
    from . import stepctx
    sc = stepctx.current()
    xchg = sc.xchg if sc is not None else None
    return (step_state, xchg, sc.next_xchg_offset() if xchg is not None else 0)


def uaps_unlabeled_loss(logits: Sequence[torch.Tensor], mix_w: Optional[Sequence[float]], cw1: float, cw2: float, *,
                        group=None, return_pseudo: bool = False, return_exp_var: bool = False,
                        exact_math: bool = False, n_global: Optional[int] = None, step_state=None):
    """Unlabeled-batch loss of UAPS for K decoders (UAPS_train.py:186-189, 223-282).

    logits: K tensors [B, C, H, W] fp32 (main, aux1, ...); mix_w: the Dirichlet draw of :251 (K floats);
    cw1 / cw2: the ramped consistency weights of :279-280.
    group: a torch.distributed process group whose ranks each hold a shard of the unlabeled batch --
    the Dice/CE/mean(exp(-KL)) sums are then taken over the WHOLE batch, as the reference's
    DataParallel gather does (UAPS_model.py:13), via one all-reduce of a <=70-double vector.

    Returns ``(loss_u, ps_loss, l_uncert, pseudo_label | None, exp_var | None)`` with
    loss_u = cw1 * ps_loss + cw2 * l_uncert (the unlabeled terms of :282), pseudo_label int64 [B,H,W]
    (bit-exact torch.argmax of the mix), exp_var a list of K [B,H,W] maps exp(-KL_k).

    step_state: a ``stepctx.DeviceStepState`` -- mix_w, cw1 and cw2 are then read from device memory (whatever is
    passed for them is ignored), which is what lets ``UAPSTrainer`` capture the iteration into a CUDA graph.
    """
    flags = L.LOSS_EXACT if exact_math else 0
    out = _FusedLossFn.apply(None if step_state is not None else tuple(float(w) for w in mix_w),
                             0.0 if step_state is not None else cw1, 0.0 if step_state is not None else cw2, None, group,
                             return_pseudo, return_exp_var, flags, n_global, _dev_ctx(step_state), *logits)
    sc = out[0]
    pseudo, exp_var = _split_extra(out, return_pseudo, return_exp_var)
    return sc[L.SC_LOSS_U], sc[L.SC_PS_LOSS], sc[L.SC_L_UNCERT], pseudo, exp_var


def uaps_unlabeled_loss_terms(logits, mix_w, cw1, cw2, **kw):
    """Like ``uaps_unlabeled_loss`` but also returns the per-decoder scalars the loop may log:
    dict(ps_k, ebar_k, ce_k, dice_k) as detached [K] tensors."""
    flags = L.LOSS_EXACT if kw.pop("exact_math", False) else 0
    out = _FusedLossFn.apply(tuple(float(w) for w in mix_w), cw1, cw2, None, kw.get("group"), False, False,
                             flags, kw.get("n_global"), None, *logits)
    K = len(logits)
    sc, b = out[0], L.SC_BASE
    d = sc.detach()
    return sc[L.SC_LOSS_U], sc[L.SC_PS_LOSS], sc[L.SC_L_UNCERT], {
        "ps_k": d[b:b + K], "ebar_k": d[b + K:b + 2 * K], "ce_k": d[b + 2 * K:b + 3 * K],
        "dice_k": d[b + 3 * K:b + 4 * K]}


def uaps_supervised_loss(logits: Sequence[torch.Tensor], labels: torch.Tensor, *, group=None,
                         exact_math: bool = False, n_global: Optional[int] = None, step_state=None):
    """Labeled-batch loss (UAPS_train.py:194-218): returns (supervised_loss, total_loss_ce, total_loss_dice,
    ce_k[K]) where supervised_loss = mean_k 0.5 (CE_k + Dice_k); all three are differentiable."""
    flags = L.LOSS_EXACT if exact_math else 0
    out = _FusedLossFn.apply(None, 1.0, 0.0, labels, group, False, False, flags, n_global, _dev_ctx(step_state), *logits)
    K = len(logits)
    sc, b = out[0], L.SC_BASE
    return sc[L.SC_LOSS_U], sc[L.SC_MEAN_CE], sc[L.SC_MEAN_DICE], sc.detach()[b + 2 * K:b + 3 * K]


def dice_loss(true: torch.Tensor, logits: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """Drop-in for utilities/pytorch_losses.py:54 ``dice_loss(true[B,1,H,W], logits[B,C,H,W])``."""
    if eps != 1e-7:
        raise RuntimeError("the fused kernel implements the reference's eps = 1e-7 only")
    out = _FusedLossFn.apply(None, 1.0, 0.0, true.squeeze(1).long(), None, False, False, 0, None, None, logits)
    return out[0][L.SC_MEAN_DICE]


def ce_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Drop-in for the reference's ``CrossEntropyLoss()(logits, target[B,H,W])`` (UAPS_train.py:75)."""
    out = _FusedLossFn.apply(None, 1.0, 0.0, target.long(), None, False, False, 0, None, None, logits)
    return out[0][L.SC_MEAN_CE]
