"""PID Lagrangian multiplier (host scalar; a8 in SURVEY.md section 8).

Same update rule and state_dict keys as the reference's ``LagrangianOptimizer``
(/root/reference/fsrl/utils/optim_util.py:20-62) so checkpoints interchange; the multiplier
is consumed by the fused loss kernels as a launch argument.
"""
from __future__ import annotations

import numpy as np


class LagrangianOptimizer(object):
    """lambda <- max(0, Kp*e + Ki*I + Kd*d) with e = mean(value - threshold),
    I <- max(0, I + e), d = max(0, e - e_prev)   (Stooke et al. 2020)."""

    def __init__(self, pid: tuple = (0.05, 0.0005, 0.1)) -> None:
        assert len(pid) == 3, " the pid param should be a list with 3 numbers"
        self.pid = tuple(pid)
        self.error_old = 0.0
        self.error_integral = 0.0
        self.lagrangian = 0.0

    def step(self, value, threshold) -> None:
        kp, ki, kd = self.pid
        err = np.mean(value - threshold)
        rise = max(0.0, err - self.error_old)
        self.error_integral = max(0.0, self.error_integral + err)
        self.error_old = err
        self.lagrangian = max(0.0, kp * err + ki * self.error_integral + kd * rise)

    def get_lag(self) -> float:
        return self.lagrangian

    def state_dict(self) -> dict:
        return {"pid": self.pid, "error_old": self.error_old,
                "error_integral": self.error_integral, "lagrangian": self.lagrangian}

    def load_state_dict(self, params: dict) -> None:
        self.pid = params["pid"]
        self.error_old = params["error_old"]
        self.error_integral = params["error_integral"]
        self.lagrangian = params["lagrangian"]
