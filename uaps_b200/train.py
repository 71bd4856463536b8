"""One UAPS training iteration (the loop body of UAPS_train.py:159-314) as a callable, data-parallel
over one process per GPU.

``UAPSTrainer.step(x_l, y_l, x_u)`` does what the reference does per iteration -- labeled forward
(:177), unlabeled forward (:185), supervised CE+Dice over the K decoders (:194-218), the fused
pseudo-label / KL-uncertainty / weighted CE+Dice loss (:223-277), the sigmoid ramp (:279-280), the
total (:282), ``zero_grad / backward / Adam.step`` (:285-292) -- and returns the logged scalars as
device tensors (the reference's nine ``.item()`` host syncs per iteration, :295-306, are left to
the caller, once per epoch).

Multi-GPU (replaces ``nn.DataParallel``, UAPS_model.py:13): each rank holds B/G labeled + B/G
unlabeled images and a full parameter replica.  The loss partial sums are all-reduced inside the
loss functions so CE/Dice/mean(exp(-KL)) are taken over the WHOLE batch like the reference's
gather-to-GPU-0 does; per-pixel gradients are therefore already gradients of the global loss and
parameter gradients are SUM-all-reduced in one flat NCCL call.  BatchNorm statistics stay per
rank, which is what DataParallel's per-replica BN does.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from .conv import pack_scope
from .stepctx import StepContext
from .losses import uaps_supervised_loss, uaps_unlabeled_loss
from .ramps import get_current_consistency_weight


@dataclass
class UAPSConfig:
    """The reference's argparse / module constants (UAPS_train.py:36-60, 66, 78)."""
    num_classes: int = 4
    base_lr: float = 1e-3
    consistency1: float = 0.1
    consistency2: float = 0.1
    consistency_rampup: float = 200.0
    iters_per_ramp_epoch: int = 80       # ``iter_num // 80`` (:279-280); 60/40/50 in the dataset variants
    seed: int = 1337
    optimizer: str = "uaps"              # "uaps": FlatAdam (one kernel over flat buffers); "torch": torch.optim.Adam(fused=True)


class FlatGradBuffer:
    """All parameter gradients as views of one contiguous fp32 buffer: zeroed with one memset,
    all-reduced with one NCCL call (14.9 MB for the 3.71 M parameters of UNet_UAPS).  Every parameter's slot starts
    on a 256-byte boundary (the kernels take 16-byte vector loads of biases and BN parameters)."""

    ALIGN = 64                            # elements: every parameter starts on a 256-byte boundary, like a torch allocation

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def zero(self):
        self.flat.zero_()

    def all_reduce_sum(self, group=None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


class FlatAdam:
    """torch.optim.Adam(params, lr) (UAPS_train.py:112) as ONE kernel over flat buffers (uaps_adam_step).

    The parameters are re-homed into one contiguous fp32 buffer (each ``p.data`` becomes a view of it, like the
    ``.grad`` views of ``FlatGradBuffer``), so the update of all 3.71 M parameters is a single streaming launch.
    ``state_dict()`` / ``load_state_dict()`` speak torch.optim.Adam's format (state[i] = step / exp_avg / exp_avg_sq,
    one param group), so the reference's checkpoints (:443-450) round-trip through either optimizer."""

    def __init__(self, grads: "FlatGradBuffer", lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.grads, self.params = grads, grads.params
        self.lr, self.betas, self.eps, self.step_count = float(lr), (float(betas[0]), float(betas[1])), float(eps), 0
        dev = grads.flat.device
        self.flat_p = torch.zeros_like(grads.flat)     # same layout as the gradients (256-byte aligned slots, zero padding)
        with torch.no_grad():
            for p, off in zip(self.params, grads.offsets):
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view                          # parameter storage now lives in the flat buffer
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.param_groups = [{"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                              "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                              "fused": None, "params": list(range(len(self.params)))}]
        assert dev.type == "cuda", "uaps_b200 has no CPU path"

    @torch.no_grad()
    def step(self):
        from . import _lib as L
        self.step_count += 1
        lr = float(self.param_groups[0]["lr"])          # a scheduler (ReduceLROnPlateau, :113) edits the group in place
        with L.on_device(self.flat_p.device):
            L.check(L.lib().uaps_adam_step(self.flat_p.data_ptr(), self.grads.flat.data_ptr(), self.exp_avg.data_ptr(),
                                           self.exp_avg_sq.data_ptr(), self.flat_p.numel(), self.step_count, lr,
                                           self.betas[0], self.betas[1], self.eps, 1.0, L.stream_ptr()), "uaps_adam_step")

    def zero_grad(self, set_to_none: bool = False):
        self.grads.zero()

    def state_dict(self):
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.grads.offsets)):
            n = p.numel()
            if self.step_count > 0:
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
        return {"state": state, "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        steps = set()
        for i, (p, off) in enumerate(zip(self.params, self.grads.offsets)):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + n].view_as(p).copy_(st["exp_avg"])
                self.exp_avg_sq[off:off + n].view_as(p).copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise RuntimeError("FlatAdam keeps one step count for all parameters; the checkpoint has several")
        self.step_count = steps.pop() if steps else 0
        g = sd["param_groups"][0]
        self.param_groups[0].update({k: g[k] for k in ("lr", "betas", "eps") if k in g})
        self.betas, self.eps = tuple(float(b) for b in self.param_groups[0]["betas"]), float(self.param_groups[0]["eps"])


class UAPSTrainer:
    def __init__(self, model: torch.nn.Module, cfg: Optional[UAPSConfig] = None, group=None):
        if group is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            group = dist.group.WORLD             # one process per GPU: the loss sums and the gradients span all ranks
        self.model, self.cfg, self.group = model, cfg or UAPSConfig(), group
        self.grads = FlatGradBuffer(model.parameters())
        if self.cfg.optimizer == "torch":
            self.optimizer = torch.optim.Adam(self.grads.params, lr=self.cfg.base_lr, fused=True)   # :112 (library kernel)
        else:
            self.optimizer = FlatAdam(self.grads, lr=self.cfg.base_lr)                              # :112, one own kernel
        self.iter_num = 0
        # identical Dirichlet draws on every rank (the reference draws once per iteration on the host, :251)
        self.rng = np.random.default_rng(self.cfg.seed)
        self.k = len(model.decoders()) if hasattr(model, "decoders") else 4

    def consistency_weights(self):
        c = self.cfg
        return (get_current_consistency_weight(self.iter_num, c.consistency1, c.consistency_rampup, c.iters_per_ramp_epoch),
                get_current_consistency_weight(self.iter_num, c.consistency2, c.consistency_rampup, c.iters_per_ramp_epoch))

    def step(self, x_l: torch.Tensor, y_l: torch.Tensor, x_u: torch.Tensor,
             mix_w=None, rand_l=None, rand_u=None) -> Dict[str, torch.Tensor]:
        self.model.train()
        with pack_scope(), StepContext(x_l.device):          # weights packed once; gradients accumulated in place
            out_l = self.model(x_l) if rand_l is None else self.model(x_l, rand=rand_l)         # :177
            out_u = self.model(x_u) if rand_u is None else self.model(x_u, rand=rand_u)         # :185
            sup, tce, tdice, ce_k = uaps_supervised_loss(out_l, y_l, group=self.group)           # :194-218
            if mix_w is None:
                mix_w = self.rng.dirichlet(np.ones(self.k))                                      # :251
            cw1, cw2 = self.consistency_weights()                                                # :279-280
            loss_u, ps_loss, l_unc, _, _ = uaps_unlabeled_loss(out_u, mix_w, cw1, cw2, group=self.group)  # :223-277
            loss = sup + loss_u                                                                  # :282
            self.grads.zero()                                                                    # :285
            loss.backward()                                                                      # :287
        self.grads.all_reduce_sum(self.group)
        self.optimizer.step()                                                                # :292
        self.iter_num += 1
        return {"loss": loss.detach(), "supervised_loss": sup.detach(), "total_loss_ce": tce.detach(),
                "total_loss_dice": tdice.detach(), "ps_loss": ps_loss.detach(), "l_uncert": l_unc.detach(),
                "loss_ce_k": ce_k}
