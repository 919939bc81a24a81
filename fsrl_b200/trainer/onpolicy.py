"""On-policy trainer (reference: /root/reference/fsrl/trainer/onpolicy.py:92-109)."""
from __future__ import annotations

from typing import Any, Dict

from .base_trainer import BaseTrainer


class OnpolicyTrainer(BaseTrainer):
    def __init__(self, policy, train_collector, test_collector=None, max_epoch: int = 10000,
                 batch_size: int = 512, cost_limit: float = float("inf"),
                 step_per_epoch: int = 10000, repeat_per_collect: int = 4,
                 episode_per_collect: int = 10, save_model_interval: int = 1,
                 episode_per_test=None, stop_fn=None, resume_from_log: bool = False,
                 logger=None, verbose: bool = True, show_progress: bool = True):
        from ..utils.logger import DummyLogger
        super().__init__("onpolicy", policy, train_collector, test_collector, max_epoch, batch_size,
                         cost_limit, step_per_epoch, repeat_per_collect, 1, save_model_interval,
                         episode_per_test, episode_per_collect, stop_fn, resume_from_log,
                         logger if logger is not None else DummyLogger(), verbose, show_progress)

    def policy_update_fn(self, stats_train: Dict[str, Any]) -> None:
        assert self.train_collector is not None
        buf = self.train_collector.buffer
        self.policy.pre_update_fn(stats_train=stats_train, batch_size=self.batch_size, buffer=buf)
        # sample_size 0: the whole buffer, sub-buffer by sub-buffer (base_policy.py:348)
        self.policy.update(0, buf, batch_size=self.batch_size, repeat=self.repeat_per_collect)
        self.policy.post_update_fn(stats_train=stats_train)
        self.train_collector.reset_buffer(keep_statistics=True)


def onpolicy_trainer(*args, **kwargs):
    return OnpolicyTrainer(*args, **kwargs).run()
