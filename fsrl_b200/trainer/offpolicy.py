"""Off-policy trainer (reference: /root/reference/fsrl/trainer/offpolicy.py:93-106)."""
from __future__ import annotations

from typing import Any, Dict

from .base_trainer import BaseTrainer


class OffpolicyTrainer(BaseTrainer):
    def __init__(self, policy, train_collector, test_collector=None, max_epoch: int = 1000,
                 batch_size: int = 512, cost_limit: float = float("inf"),
                 step_per_epoch: int = 10000, update_per_step: float = 0.1,
                 episode_per_collect: int = 1, save_model_interval: int = 1,
                 episode_per_test=None, stop_fn=None, resume_from_log: bool = False,
                 logger=None, verbose: bool = True, show_progress: bool = True):
        from ..utils.logger import DummyLogger
        super().__init__("offpolicy", policy, train_collector, test_collector, max_epoch, batch_size,
                         cost_limit, step_per_epoch, None, update_per_step, save_model_interval,
                         episode_per_test, episode_per_collect, stop_fn, resume_from_log,
                         logger if logger is not None else DummyLogger(), verbose, show_progress)
        self.gradient_steps = 0

    def policy_update_fn(self, stats_train: Dict[str, Any]) -> None:
        assert self.train_collector is not None
        buf = self.train_collector.buffer
        self.policy.pre_update_fn(stats_train=stats_train, batch_size=self.batch_size, buffer=buf,
                                  update_per_step=self.update_per_step)
        # data-parallel ranks derive the number of gradient steps from the step count they agreed on
        # (BaseTrainer._agreed_steps), so every rank joins the same number of gradient exchanges
        n_updates = round(self.update_per_step * getattr(self, "_cycle_steps", stats_train["n/st"]))
        if hasattr(self.policy, "update_many"):
            # same gradient steps, launched back to back without returning to Python each time
            self.policy.update_many(n_updates, self.batch_size, buf)
            self.gradient_steps += n_updates
        else:
            for _ in range(n_updates):
                self.gradient_steps += 1
                self.policy.update(self.batch_size, buf)
        self.policy.post_update_fn(stats_train=stats_train)
        self.logger.store(gradient_steps=self.gradient_steps, tab="update")


def offpolicy_trainer(*args, **kwargs):
    return OffpolicyTrainer(*args, **kwargs).run()
