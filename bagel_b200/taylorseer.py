"""Host-side schedule of the TaylorSeer step cache (reference modeling/cache_utils/taylorseer.py), one instance per
velocity branch as in Bagel.generate_image (reference bagel.py:680-684). Only integers live here; the factor planes
and the two kernels that touch them (ops.taylor_update / ops.taylor_eval) are on the device.

Reference behaviour reproduced (`cache_init` :128-166 -> fresh_threshold 3, max_order 6, first_enhance 5;
`cal_type` :80-122; `force_scheduler` :62-76 with linear_step_weight 0 -> threshold 3):
  * steps 0..4 are fully computed, afterwards every 3rd step (7, 10, 13, ...); the others are extrapolated;
  * a fully computed step refreshes the factors: order k+1 exists once order k existed at the previous full step and
    `step > first_enhance - 2` (:23), capped at max_order;
  * an extrapolated step evaluates the Taylor polynomial at x = step - last_full_step (:40).
"""
from __future__ import annotations

from typing import Tuple


class TaylorSeerSchedule:
    FRESH_THRESHOLD = 3
    MAX_ORDER = 6
    FIRST_ENHANCE = 5

    def __init__(self, num_steps: int):
        self.num_steps = num_steps
        self.step = 0
        self.cache_counter = 0
        self.cal_threshold = None
        self.activated_steps = [0]
        self.n_factors = 0          # factor planes currently valid
        self.type = None

    def begin_step(self) -> str:
        """cal_type: decide 'full' or 'Taylor' for the evaluation about to run."""
        first = self.step < self.FIRST_ENHANCE
        interval = self.FRESH_THRESHOLD if first else self.cal_threshold
        if first or self.cache_counter == interval - 1:
            self.type = "full"
            self.cache_counter = 0
            self.activated_steps.append(self.step)
            self.cal_threshold = int(round(self.FRESH_THRESHOLD / 1.0))
        else:
            self.cache_counter += 1
            self.type = "Taylor"
        return self.type

    def full_update_args(self) -> Tuple[int, int]:
        """(n_deriv, dist) for ops.taylor_update after a fully computed step; updates the valid-plane count."""
        assert self.type == "full"
        if self.step == 0:
            self.n_factors = 0      # taylor_cache_init (:49-58)
        dist = self.activated_steps[-1] - self.activated_steps[-2]
        n_deriv = min(self.n_factors, self.MAX_ORDER) if self.step > self.FIRST_ENHANCE - 2 else 0
        self.n_factors = n_deriv + 1
        return n_deriv, dist

    def taylor_args(self) -> Tuple[int, int]:
        """(n_factors, x) for ops.taylor_eval on an extrapolated step."""
        assert self.type == "Taylor"
        return self.n_factors, self.step - self.activated_steps[-1]

    def end_step(self):
        self.step += 1
