"""Per-iteration context of the bf16 training path: things that are safe only because one training
iteration runs between a gradient-buffer zeroing and an optimizer step (``UAPSTrainer.step`` opens it).

* ``direct_grads``: weight / BatchNorm gradients are accumulated by the kernels straight into the
  parameters' pre-zeroed ``.grad`` views (no per-parameter zeros + AccumulateGrad add launches).
* a zero-filled fp64 arena for the BatchNorm statistic sums (one memset per iteration instead of one per layer).
* deferred ``num_batches_tracked`` increments (one foreach add per iteration).
"""
from __future__ import annotations

import torch

_active = None


class StepContext:
    def __init__(self, device, arena_doubles: int = 1 << 17):
        self.device = device
        self.arena = torch.zeros(arena_doubles, dtype=torch.float64, device=device)
        self.used = 0
        self.counters = []
        self.direct_grads = True
        self.stream = None                 # raw cudaStream_t of the iteration, cached for the ~1150 launches (see _lib.stream_ptr)

    def take(self, n: int) -> torch.Tensor:
        """n zeroed doubles (falls back to a fresh tensor when the arena is exhausted)."""
        if self.used + n > self.arena.numel():
            return torch.zeros(n, dtype=torch.float64, device=self.device)
        out = self.arena[self.used:self.used + n]
        self.used += n
        return out

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
