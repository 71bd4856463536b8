"""Host-side ramp schedule (utilities/ramps.py:19-26; UAPS_train.py:81-87, 279-280). Pure scalar math."""
from __future__ import annotations

import math


def sigmoid_rampup(current: float, rampup_length: float) -> float:
    """exp(-5 (1 - clip(t, 0, T) / T)^2); 1.0 when T == 0  (utilities/ramps.py:19-26)."""
    if rampup_length == 0:
        return 1.0
    cur = min(max(float(current), 0.0), float(rampup_length))
    phase = 1.0 - cur / rampup_length
    return float(math.exp(-5.0 * phase * phase))


def get_current_consistency_weight(iter_num: int, consistency: float = 0.1, consistency_rampup: float = 200.0,
                                   iters_per_ramp_epoch: int = 80) -> float:
    """consistency * sigmoid_rampup(iter_num // 80, rampup)  (UAPS_train.py:81-87 called at :279-280;
    the DAGM / KoSDD2 / MTiles variants divide by 60 / 40 / 50)."""
    return consistency * sigmoid_rampup(iter_num // iters_per_ramp_epoch, consistency_rampup)
