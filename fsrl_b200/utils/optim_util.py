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


class RunningMeanStd:
    """Running mean / variance of a data stream (tianshou.utils.RunningMeanStd, the statistic behind
    ``reward_normalization`` in fsrl/policy/base_policy.py:111,434-444): parallel-variance update
    of (mean, var, count) per call.  [tianshou 0.5.0 source absent: restated from its documented
    behaviour, SURVEY.md 2.3]"""

    def __init__(self, mean=0.0, std=1.0, clip_max=10.0, epsilon=float(np.finfo(np.float32).eps)):
        self.mean, self.var = mean, std
        self.clip_max = clip_max
        self.count = 0
        self.eps = epsilon

    def update_moments(self, batch_mean, batch_var, batch_count) -> None:
        delta = batch_mean - self.mean
        total = self.count + batch_count
        m2 = self.var * self.count + batch_var * batch_count + delta ** 2 * self.count * batch_count / total
        self.mean = self.mean + delta * batch_count / total
        self.var = m2 / total
        self.count = total

    def update(self, x) -> None:
        x = np.asarray(x)
        self.update_moments(np.mean(x, axis=0), np.var(x, axis=0), len(x))
