"""Oracle (test infrastructure): PID Lagrangian multiplier.

Restates /root/reference/fsrl/utils/optim_util.py:28-45 (LagrangianOptimizer.step /
get_lag) and the driver /root/reference/fsrl/policy/lagrangian_base.py:98-120,145-166.
Pinned by tests/golden/pid_golden.json (produced by importing the reference class).
"""
from __future__ import annotations

import numpy as np


class PIDLagrangian:
    def __init__(self, pid=(0.05, 0.0005, 0.1)):
        assert len(pid) == 3
        self.kp, self.ki, self.kd = (float(x) for x in pid)
        self.error_old = 0.0
        self.error_integral = 0.0
        self.lagrangian = 0.0

    def step(self, value, threshold):
        e = float(np.mean(np.asarray(value, dtype=np.float64) - threshold))   # :34
        d = max(0.0, e - self.error_old)                                      # :35
        self.error_integral = max(0.0, self.error_integral + e)               # :36
        self.error_old = e                                                    # :37
        self.lagrangian = max(0.0, self.kp * e + self.ki * self.error_integral
                              + self.kd * d)                                  # :38-41
        return self.lagrangian


def rescaling_factor(lags, rescaling=True):
    """lagrangian_base.py:156."""
    return 1.0 / (float(np.sum(lags)) + 1.0) if rescaling else 1.0
