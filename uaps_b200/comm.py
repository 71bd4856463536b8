"""Peer-memory exchange for the loss sums: the host half of uaps_xchg_* / uaps_loss_pass1_exchange.

One ``LossExchange`` per (process group, device): allocates this rank's mailbox, trades CUDA-IPC handles with
the other ranks through ``torch.distributed`` (plumbing only -- no collective runs on the data path afterwards),
maps every peer's mailbox and hands the fused kernel the pointer table plus a per-call epoch.  All ranks must
call the loss functions in the same order (they do: one training loop per rank), which keeps the epochs equal.

Used when every rank of the group sits on this node and UAPS_LOSS_EXCHANGE is not "nccl"; otherwise the loss
falls back to fold -> ncclAllReduce -> finalize (still device code, but three launches and NCCL's latency).
"""
from __future__ import annotations

import ctypes as C
import os
import socket
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib as L

MAX_RANKS = 8


class LossExchange:
    def __init__(self, group=None, device: Optional[torch.device] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > MAX_RANKS:
            raise RuntimeError(f"the peer-memory exchange supports up to {MAX_RANKS} ranks, got {self.world}")
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = L.lib()
        self.own, self._opened, self.ptrs = None, [], (C.c_void_p * self.world)()
        # Every rank runs every collective below whatever happens locally (a rank that bailed out early would hang
        # the others); failures are carried as None / flags and raised together at the end.
        with torch.cuda.device(self.device):
            mine = None
            own = C.c_void_p()
            if lib.uaps_xchg_alloc(C.byref(own)) == 0:
                self.own = own.value
                handle = C.create_string_buffer(64)
                if lib.uaps_xchg_export(self.own, handle) == 0:
                    mine = bytes(handle.raw)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            failed = any(h is None for h in handles)
            if not failed:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self.ptrs[r] = self.own
                        continue
                    p = C.c_void_p()
                    if lib.uaps_xchg_import(C.create_string_buffer(h, 64), C.byref(p)) != 0:
                        failed = True
                        break
                    self.ptrs[r] = p.value
                    self._opened.append(p.value)
        if failed:
            self.close()
            raise RuntimeError("could not allocate / export / map the exchange mailboxes (CUDA IPC or peer access unavailable)")
        self.epoch = 0

    def next_epoch(self) -> int:
        # 1, 2, 3, ... as a 32-bit tag; the wrap skips 0 (the "never written" tag) and keeps the parity alternating
        self.epoch = self.epoch + 1 if self.epoch < 0xFFFFFFFF else 2
        return self.epoch

    def status(self) -> int:
        """0, or the epoch of the first exchange that timed out waiting for a peer (synchronises the stream)."""
        out = C.c_uint(0)
        with torch.cuda.device(self.device):
            L.check(L.lib().uaps_xchg_status(self.own, C.byref(out), L.stream_ptr()), "uaps_xchg_status")
        return int(out.value)

    def close(self) -> None:
        lib = L.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in getattr(self, "_opened", []):
                lib.uaps_xchg_close(p)
            self._opened = []
            if self.own:
                lib.uaps_xchg_free(self.own)
                self.own = None


class PeerRegion:
    """`nbytes` of zero-filled device memory on every rank of `group`, each mapped into every other rank (CUDA IPC).
    ``ptrs[r]`` is rank r's buffer as seen from this process (own at ``ptrs[rank]``).  COLLECTIVE constructor; raises
    on every rank together if any rank could not allocate / export / map."""

    def __init__(self, group, device: torch.device, nbytes: int):
        self.world, self.rank, self.device, self.nbytes = dist.get_world_size(group), dist.get_rank(group), device, int(nbytes)
        lib = L.lib()
        self.own, self._opened, self.ptrs = None, [], (C.c_void_p * self.world)()
        with torch.cuda.device(device):
            mine, own = None, C.c_void_p()
            if lib.uaps_peer_alloc(C.byref(own), self.nbytes) == 0:
                self.own = own.value
                handle = C.create_string_buffer(64)
                if lib.uaps_xchg_export(self.own, handle) == 0:
                    mine = bytes(handle.raw)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            failed = any(h is None for h in handles)
            if not failed:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self.ptrs[r] = self.own
                        continue
                    p = C.c_void_p()
                    if lib.uaps_xchg_import(C.create_string_buffer(h, 64), C.byref(p)) != 0:
                        failed = True
                        break
                    self.ptrs[r] = p.value
                    self._opened.append(p.value)
            outcomes = [None] * self.world
            dist.all_gather_object(outcomes, failed, group=group)        # agree on the outcome; also: everything is mapped
        if any(outcomes):
            self.close()
            raise RuntimeError("could not allocate / export / map a peer region (CUDA IPC or peer access unavailable)")

    def tensor(self, dtype=torch.float32) -> torch.Tensor:
        """The own buffer as a torch tensor (no copy; this object keeps the allocation alive and must outlive the tensor)."""
        itemsize = torch.empty((), dtype=dtype).element_size()
        typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]
        region = self

        class _Arr:
            __cuda_array_interface__ = {"shape": (region.nbytes // itemsize,), "typestr": typestr, "data": (region.own, False),
                                        "version": 2, "strides": None}
        holder = _Arr()
        t = torch.as_tensor(holder, device=self.device)
        t._uaps_region = self                      # keep the allocation alive as long as the tensor object lives
        return t

    def close(self) -> None:
        lib = L.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in getattr(self, "_opened", []):
                lib.uaps_xchg_close(p)
            self._opened = []
            if self.own:
                lib.uaps_peer_free(self.own)
                self.own = None


class PeerGradReducer:
    """Host half of ``uaps_grad_reduce_adam``: a peer-mapped flat gradient buffer of `numel` floats on every rank plus the
    flag boxes of the two barriers.  ``UAPSTrainer`` builds its ``FlatGradBuffer`` on ``grad_tensor()`` so that the
    backward kernels write the gradients straight into memory the peers can read."""

    def __init__(self, group, device: torch.device, numel: int):
        self.grads = PeerRegion(group, device, 4 * int(numel))
        self.flags = PeerRegion(group, device, L.lib().uaps_xchg_mailbox_bytes())
        self.rank, self.world, self.numel = self.grads.rank, self.grads.world, int(numel)

    def grad_tensor(self) -> torch.Tensor:
        return self.grads.tensor(torch.float32)

    def status(self) -> int:
        """0, or the epoch of the first reduce that timed out waiting for a peer (synchronises)."""
        return int(self.flags.tensor(torch.int32)[81].item())

    def close(self) -> None:
        self.grads.close()
        self.flags.close()


def new_grad_reducer(group, device: torch.device, numel: int) -> Optional[PeerGradReducer]:
    """COLLECTIVE.  None when peer memory cannot be used (same conditions as ``new_exchange``)."""
    if not dist.is_initialized():
        return None
    ok = (os.environ.get("UAPS_GRAD_REDUCE", "peer") != "nccl" and dist.get_backend(group) == "nccl"
          and 1 < dist.get_world_size(group) <= MAX_RANKS)
    if ok:
        hosts = [None] * dist.get_world_size(group)
        dist.all_gather_object(hosts, socket.gethostname(), group=group)
        ok = len(set(hosts)) == 1
    if not ok:
        return None
    try:
        red = PeerGradReducer(group, device, numel)
    except RuntimeError as e:               # raised on every rank together (PeerRegion agrees on the outcome collectively)
        if dist.get_rank(group) == 0:
            import warnings
            warnings.warn(f"uaps_b200: peer-memory gradient reduce unavailable, using the NCCL all-reduce: {e}")
        return None
    _owned.append(red)
    return red


_exchanges: Dict[tuple, Optional[LossExchange]] = {}
_owned = []


def new_exchange(group, device: torch.device) -> Optional[LossExchange]:
    """A fresh exchange (own mailboxes, epochs starting at 0) for (group, device), or None when the group cannot use
    peer memory (not NCCL, ranks on several hosts, more than 8 ranks, UAPS_LOSS_EXCHANGE=nccl, or the mapping failed on
    any rank).  COLLECTIVE: every rank of the group must call it, and every rank gets the same kind of answer.
    ``UAPSTrainer`` takes one per trainer: its exchange epochs are counted on the device (UapsStepState.xchg_base), so
    they must not share mailboxes with the host-counted epochs of direct ``uaps_unlabeled_loss(group=...)`` calls."""
    if not dist.is_initialized():
        return None
    ok = (os.environ.get("UAPS_LOSS_EXCHANGE", "peer") != "nccl" and dist.get_backend(group) == "nccl"
          and 1 < dist.get_world_size(group) <= MAX_RANKS)
    if ok:                                   # collective decision: every rank must take the same branch
        hosts = [None] * dist.get_world_size(group)
        dist.all_gather_object(hosts, socket.gethostname(), group=group)
        ok = len(set(hosts)) == 1
    if not ok:
        return None
    # mapping a peer's mailbox can fail (no P2P path, IPC disabled in the container ...): every rank must then
    # take the NCCL route, so the outcome is agreed on collectively before anybody uses the exchange
    xchg, err = None, None
    try:
        xchg = LossExchange(group, device)
    except Exception as e:                 # noqa: BLE001 -- reported below, then the NCCL fallback is used
        err = repr(e)
    outcomes = [None] * dist.get_world_size(group)
    dist.all_gather_object(outcomes, err, group=group)     # also the barrier: every mailbox zeroed and mapped
    if any(o is not None for o in outcomes):
        if xchg is not None:
            xchg.close()
        xchg = None
        if dist.get_rank(group) == 0:
            import warnings
            warnings.warn("uaps_b200: peer-memory exchange unavailable, using the NCCL all-reduce for the loss sums: "
                          + "; ".join(o for o in outcomes if o is not None)[:300])
    if xchg is not None:
        _owned.append(xchg)
    return xchg


def exchange_for(group, device: torch.device) -> Optional[LossExchange]:
    """The shared, host-epoch exchange of (group, device), created on first use (see ``new_exchange`` for None)."""
    if not dist.is_initialized():
        return None
    key = (id(group) if group is not None else 0, device.index)
    if key not in _exchanges:
        _exchanges[key] = new_exchange(group, device)
    return _exchanges[key]


def close_all() -> None:
    for x in _owned:
        x.close()
    _owned.clear()
    _exchanges.clear()
