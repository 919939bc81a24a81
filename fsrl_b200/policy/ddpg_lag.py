"""DDPG with PID Lagrangian (reference: /root/reference/fsrl/policy/ddpg_lag.py): deterministic
tanh actor with Gaussian exploration noise (added by the rollout kernel), one Q-critic per
return stream, n-step targets from the target actor/critics, Polyak sync of actor + critics."""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..data.batch import Batch
from ..nets import Actor
from ..utils.logger import BaseLogger, DummyLogger
from .offpolicy_base import OffPolicyLagrangian


class GaussianNoise:
    """tianshou.exploration.GaussianNoise surface: only sigma is consumed (by the rollout kernel)."""

    def __init__(self, mu: float = 0.0, sigma: float = 1.0):
        self._mu, self._sigma = mu, sigma
        assert 0 <= sigma, "Noise std should not be negative."

    def __call__(self, size):
        return np.random.normal(self._mu, self._sigma, size)

    def reset(self):
        pass


class DDPGLagrangian(OffPolicyLagrangian):
    _algo = _lib.ALGO_DDPG

    def __init__(self, actor, critics, actor_optim, critic_optim, logger: BaseLogger = DummyLogger(),
                 tau: float = 0.05, exploration_noise: Optional[GaussianNoise] = GaussianNoise(sigma=0.1),
                 n_step: int = 2, use_lagrangian: bool = True,
                 lagrangian_pid: Tuple = (0.05, 0.0005, 0.1), cost_limit=np.inf, rescaling: bool = True,
                 gamma: float = 0.99, reward_normalization: bool = False, deterministic_eval: bool = True,
                 action_scaling: bool = True, action_bound_method: str = "clip", observation_space=None,
                 action_space=None, lr_scheduler=None) -> None:
        super().__init__(actor, critics, None, logger, use_lagrangian, lagrangian_pid, cost_limit,
                         rescaling, gamma, 99999, reward_normalization, deterministic_eval,
                         action_scaling, action_bound_method, observation_space, action_space, lr_scheduler)
        if not isinstance(actor, Actor):
            raise TypeError("DDPGLagrangian needs a deterministic Actor")
        self._twin = False
        self.actor_old = deepcopy(self.actor)
        self.actor_old.eval()
        self.critics_old = deepcopy(self.critics)
        self.critics_old.eval()
        self.actor_optim, self.critics_optim = actor_optim, critic_optim
        self._init_offpolicy(tau, n_step, actor_optim.param_groups[0]["lr"], critic_optim.param_groups[0]["lr"])
        self._noise = exploration_noise

    def set_exp_noise(self, noise) -> None:
        self._noise = noise

    def _net_list(self):
        return [self.actor, self.actor_old] + list(self.critics) + list(self.critics_old)

    def _groups(self):
        g = self._slot_groups
        C = self.critics_num
        return {"actor": g[0], "actor_old": g[1],
                "critics": [s for grp in g[2:2 + C] for s in grp],
                "critics_old": [s for grp in g[2 + C:2 + 2 * C] for s in grp]}

    def fill_rollout(self, r, exploration_noise: bool = False) -> None:
        super().fill_rollout(r, exploration_noise)
        # exploration_noise (ddpg_lag.py:225-231): only when the collector asks for it
        if exploration_noise and self._noise is not None and self.training:
            # the rollout kernel adds N(0, sigma^2) itself (ddpg_lag.py:225-231 with tianshou's GaussianNoise): other
            # noise processes (OU noise, a non-zero mean, arbitrary callables) have no device twin -- refuse, don't drop
            sigma = getattr(self._noise, "_sigma", None)
            if sigma is None or float(getattr(self._noise, "_mu", 0.0)) != 0.0:
                raise TypeError("the device rollout supports zero-mean GaussianNoise(sigma) exploration only, got %r" % (self._noise,))
            r.expl_sigma = float(sigma)

    def sync_weight(self) -> None:
        g = self._groups()
        eng = self._ensure_engine(256)
        eng.polyak(g["actor_old"], g["actor"], self.tau)
        eng.polyak(g["critics_old"], g["critics"], self.tau)

    def forward(self, batch: Batch, state=None, model: str = "actor", input: str = "obs", **kwargs: Any) -> Batch:
        obs = torch.as_tensor(batch[input], dtype=torch.float32, device=self.device).contiguous()
        slot = 0 if model == "actor" else 1
        out = self.net_forward(slot, obs)
        return Batch(act=self.actor._max * torch.tanh(out), state=None)
