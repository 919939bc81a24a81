"""Soft Actor-Critic with PID Lagrangian (reference: /root/reference/fsrl/policy/sac_lag.py).
Tanh-squashed Gaussian actor (conditioned sigma), one DoubleCritic per return stream with
Polyak targets, n-step targets for reward AND cost critics (the entropy term is subtracted
from every critic's target, :141-144), optional automatic temperature."""
from __future__ import annotations

from copy import deepcopy
from typing import Any, List, Optional, Tuple, Union

import numpy as np
import torch

from .. import _lib
from ..data.batch import Batch
from ..nets import ActorProb, DoubleCritic
from ..utils.logger import BaseLogger, DummyLogger
from .offpolicy_base import OffPolicyLagrangian


class SACLagrangian(OffPolicyLagrangian):
    _algo = _lib.ALGO_SAC

    def __init__(self, actor, critics, actor_optim, critic_optim, logger: BaseLogger = DummyLogger(),
                 alpha: Union[float, Tuple] = 0.005, tau: float = 0.05, exploration_noise=None,
                 n_step: int = 2, use_lagrangian: bool = True,
                 lagrangian_pid: Tuple = (0.05, 0.0005, 0.1), cost_limit=np.inf, rescaling: bool = True,
                 gamma: float = 0.99, reward_normalization: bool = False, deterministic_eval: bool = True,
                 action_scaling: bool = True, action_bound_method: str = "clip", observation_space=None,
                 action_space=None, lr_scheduler=None) -> None:
        super().__init__(actor, critics, None, logger, use_lagrangian, lagrangian_pid, cost_limit,
                         rescaling, gamma, 99999, reward_normalization, deterministic_eval,
                         action_scaling, action_bound_method, observation_space, action_space, lr_scheduler)
        if not isinstance(actor, ActorProb) or not actor._c_sigma:
            raise TypeError("SACLagrangian needs an ActorProb with conditioned_sigma=True")
        self._twin = all(isinstance(c, DoubleCritic) for c in self.critics)
        self.critics_old = deepcopy(self.critics)
        self.critics_old.eval()
        self.actor_optim, self.critics_optim = actor_optim, critic_optim
        a_lr = actor_optim.param_groups[0]["lr"]
        c_lr = critic_optim.param_groups[0]["lr"]
        self._init_offpolicy(tau, n_step, a_lr, c_lr)
        self._is_auto_alpha = False
        if isinstance(alpha, tuple):
            self._is_auto_alpha = True
            self._target_entropy, log_alpha, alpha_optim = alpha
            self._alpha_lr = alpha_optim.param_groups[0]["lr"] if hasattr(alpha_optim, "param_groups") else float(alpha_optim)
            self._alpha0 = float(torch.as_tensor(log_alpha).detach().exp().item())
            self._log_alpha0 = float(torch.as_tensor(log_alpha).detach().item())
        else:
            self._alpha0, self._log_alpha0 = float(alpha), float(np.log(alpha)) if alpha > 0 else -np.inf
            self._alpha_lr, self._target_entropy = 0.0, 0.0
        self._noise = exploration_noise
        self._alpha_dev = None

    def set_exp_noise(self, noise) -> None:
        """sac_lag.py:125-127 (the reference's SAC agents pass None: the stochastic actor explores by itself)."""
        self._noise = noise

    def _net_list(self):
        return [self.actor] + list(self.critics) + list(self.critics_old)

    def _groups(self):
        g = self._slot_groups
        C = self.critics_num
        crit = [s for grp in g[1:1 + C] for s in grp]
        crit_old = [s for grp in g[1 + C:1 + 2 * C] for s in grp]
        return {"actor": g[0], "critics": crit, "critics_old": crit_old}

    def _fill_algo(self, d) -> None:
        if self._alpha_dev is None:
            self._alpha_dev = torch.tensor([self._alpha0], dtype=torch.float32, device=self.device)
            self._alpha_state = torch.tensor([self._log_alpha0 if np.isfinite(self._log_alpha0) else -100.0, 0.0, 0.0, 0.0],
                                             dtype=torch.float32, device=self.device)
        d.use_alpha = 1
        d.auto_alpha = int(self._is_auto_alpha)
        d.alpha_lr, d.target_entropy = float(self._alpha_lr), float(self._target_entropy)
        d.alpha, d.alpha_state = self._alpha_dev.data_ptr(), self._alpha_state.data_ptr()

    @property
    def _alpha(self):
        return float(self._alpha_dev.item()) if self._alpha_dev is not None else self._alpha0

    def _extra_stats(self, out, st):
        if self._is_auto_alpha:
            out["loss/alpha_loss"] = st[:, 5]
            out["loss/alpha_value"] = st[:, 6]

    def sync_weight(self) -> None:
        g = self._groups()
        self._ensure_engine(256).polyak(g["critics_old"], g["critics"], self.tau)

    def forward(self, batch: Batch, state=None, input: str = "obs", **kwargs: Any) -> Batch:
        """API-compatible forward (sac_lag.py:147-183) on a device batch."""
        obs = torch.as_tensor(batch[input], dtype=torch.float32, device=self.device).contiguous()
        out = self.net_forward(0, obs)
        A = out.shape[1] // 2
        from ..nets import SIGMA_MAX, SIGMA_MIN
        mu = out[:, :A]
        if not self.actor._unbounded:
            mu = self.actor._max * torch.tanh(mu)
        sigma = out[:, A:].clamp(SIGMA_MIN, SIGMA_MAX).exp()
        dist = torch.distributions.Independent(torch.distributions.Normal(mu, sigma), 1)
        act = mu if (self._deterministic_eval and not self.training) else dist.rsample()
        log_prob = dist.log_prob(act).unsqueeze(-1)
        sq = torch.tanh(act)
        log_prob = log_prob - torch.log((1 - sq.pow(2)) + np.finfo(np.float32).eps.item()).sum(-1, keepdim=True)
        return Batch(logits=(mu, sigma), act=sq, state=None, dist=dist, log_prob=log_prob)
