"""Per-iteration context of the bf16 training path: things that are safe only because one training
iteration runs between a gradient-buffer zeroing and an optimizer step (``UAPSTrainer.step`` opens it).

* ``direct_grads``: weight / BatchNorm gradients are accumulated by the kernels straight into the
  parameters' pre-zeroed ``.grad`` views (no per-parameter zeros + AccumulateGrad add launches).
* a zero-filled fp64 arena for the BatchNorm statistic sums (one memset per iteration instead of one per layer).
* deferred ``num_batches_tracked`` increments (one foreach add per iteration).
* ``state``: the device-resident per-iteration scalars (``DeviceStepState`` over ``UapsStepState``).  When it is set,
  the layers take their Philox key, FeatureDropout threshold, mix weights, ramp weights and exchange epoch from
  device memory instead of drawing them on the host, which makes the iteration a static launch sequence
  (``UAPSTrainer`` captures it into a CUDA graph).
"""
from __future__ import annotations

import ctypes as C

import torch

_active = None


class DeviceStepState:
    """One ``UapsStepState`` in device memory (include/uaps_b200.h) plus typed host access to its fields.

    The reference draws these scalars on the host every iteration (UAPS_train.py:251, 279-280; UAPS_unet.py:165);
    here ``begin()`` launches ``uaps_step_begin`` and the kernels read the struct through the pointers ``ptr()`` hands out."""

    def __init__(self, device: torch.device, lr: float):
        from . import _lib as L
        self._L = L
        self.device = device
        self.buf = torch.zeros(256, dtype=torch.uint8, device=device)       # 168 bytes used; torch allocations are 512-byte aligned
        self.base = self.buf.data_ptr()
        self._off = {n: getattr(L.StepStateStruct, n).offset for n, _ in L.StepStateStruct._fields_}
        self.set("lr", lr)

    def ptr(self, field: str, index: int = 0) -> int:
        return self.base + self._off[field] + 4 * index

    def _view(self, field: str):
        L = self._L
        ctype = dict(L.StepStateStruct._fields_)[field]
        size = C.sizeof(ctype)
        dt = torch.int64 if ctype is C.c_uint64 else (torch.int32 if ctype is C.c_uint32 else torch.float32)
        return self.buf[self._off[field]:self._off[field] + size].view(dt)

    def set(self, field: str, value) -> None:
        """Stream-ordered device write of one field (e.g. ``lr`` from a scheduler, ``iter`` / ``adam_step`` from a checkpoint)."""
        self._view(field).fill_(value)

    def read(self):
        """Host copy of the whole struct (synchronises)."""
        raw = bytes(self.buf.cpu().numpy().tobytes()[:C.sizeof(self._L.StepStateStruct)])
        return self._L.StepStateStruct.from_buffer_copy(raw)

    def begin(self, seed_rank: int, seed_shared: int, K: int, n_u: int, c1: float, c2: float, rampup: float, ipe: int,
              n_exchanges: int, beta1: float, beta2: float) -> None:
        L = self._L
        with L.on_device(self.device):
            L.check(L.lib().uaps_step_begin(self.base, seed_rank & (2 ** 64 - 1), seed_shared & (2 ** 64 - 1), K, n_u, c1, c2,
                                            rampup, ipe, n_exchanges, beta1, beta2, L.stream_ptr()), "uaps_step_begin")


class StepContext:
    def __init__(self, device, arena_doubles: int = 1 << 18, state: "DeviceStepState" = None, xchg=None, packer=None):
        self.device = device
        self.arena = torch.zeros(arena_doubles, dtype=torch.float64, device=device)
        self.used = 0
        self.counters = []
        self.direct_grads = True
        self.stream = None                 # raw cudaStream_t of the iteration, cached for the ~1150 launches (see _lib.stream_ptr)
        self.state = state                 # device-resident scalars (None: host-drawn scalars passed by value)
        self.xchg = xchg                   # this trainer's own LossExchange (device-managed epochs) or None
        self.packer = packer               # conv.WeightPacker: all layers' weights re-packed by one launch per iteration
        self._seed_calls = self._u_calls = self._xchg_calls = 0

    def take(self, n: int) -> torch.Tensor:
        """n zeroed doubles (falls back to a fresh tensor when the arena is exhausted)."""
        if self.used + n > self.arena.numel():
            return torch.zeros(n, dtype=torch.float64, device=self.device)
        out = self.arena[self.used:self.used + n]
        self.used += n
        return out

    # ---- device-state mode: per-call constants.  The call sequence of an iteration is fixed, so call i gets the
    # same constant every iteration; the per-iteration variation comes from the key uaps_step_begin wrote.
    def next_seed(self) -> int:
        self._seed_calls += 1
        return (0x9E3779B97F4A7C15 * self._seed_calls) & (2 ** 63 - 1)

    def next_u_slot(self) -> int:
        slot = self._u_calls
        self._u_calls += 1
        if slot >= 16:
            raise RuntimeError("more than 16 FeatureDropout calls in one iteration (UAPS_STEP_USLOTS)")
        return slot

    def next_xchg_offset(self) -> int:
        self._xchg_calls += 1
        return self._xchg_calls

    def __enter__(self):
        global _active
        self._prev, _active = _active, self
        if torch.cuda.is_available():
            self.stream = torch.cuda.current_stream(self.device).cuda_stream
        return self

    def __exit__(self, *exc):
        global _active
        _active = self._prev
        if self.counters:
            torch._foreach_add_(self.counters, 1)
        return False


def current():
    return _active
