"""Epoch iterator with the reference trainer's protocol and metric
(/root/reference/fsrl/trainer/base_trainer.py:100-356): ``for epoch, stats, info in trainer``;
each epoch = collect -> pre_update -> update -> post_update cycles until ``step_per_epoch``
env steps, then test, checkpoint, log.  ``train_speed`` (:345-347) is BASELINE.json's
metric: collected env steps / (wall time - test collect time).  Host control code; all the
work it drives runs on the GPU."""
from __future__ import annotations

import time
from abc import ABC, abstractmethod
from collections import deque
from typing import Any, Callable, Dict, Optional, Tuple, Union

import numpy as np

from ..utils.logger import BaseLogger, DummyLogger


class _Bar:
    """Tiny stand-in for the tqdm bar the reference drives (``t.n``, ``t.total``)."""

    def __init__(self, total, desc="", show=False):
        self.total, self.n, self.desc, self.show = total, 0, desc, show
        self._post = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        if self.show:
            print(f"{self.desc}: {self.n}/{self.total} {self._post}")
        return False

    def update(self, k):
        self.n += k

    def set_postfix(self, **kw):
        self._post = kw


class BaseTrainer(ABC):
    def __init__(self, learning_type: str, policy, train_collector, test_collector=None,
                 max_epoch: int = 100, batch_size: int = 512, cost_limit: float = np.inf,
                 step_per_epoch: Optional[int] = None, repeat_per_collect: Optional[int] = None,
                 update_per_step: Union[int, float] = 1, save_model_interval: int = 1,
                 episode_per_test: Optional[int] = None, episode_per_collect: int = 1,
                 stop_fn: Optional[Callable[[float, float], bool]] = None,
                 resume_from_log: bool = False, logger: BaseLogger = DummyLogger(),
                 verbose: bool = True, show_progress: bool = True):
        self.learning_type = learning_type
        self.policy = policy
        self.train_collector, self.test_collector = train_collector, test_collector
        self.logger = logger
        self.cost_limit = cost_limit
        self.start_time = time.time()
        self.best_perf_rew, self.best_perf_cost = -np.inf, np.inf
        self.start_epoch = 0
        self.env_step = 0
        self.cum_cost = 0
        self.cum_episode = 0
        self.max_epoch = max_epoch
        self.step_per_epoch = step_per_epoch
        self.episode_per_collect = episode_per_collect
        self.episode_per_test = episode_per_test
        self.update_per_step = update_per_step
        self.save_model_interval = save_model_interval
        self.repeat_per_collect = repeat_per_collect
        self.batch_size = batch_size
        self.stop_fn = stop_fn
        self.verbose, self.show_progress = verbose, show_progress
        self.resume_from_log = resume_from_log
        self.epoch = self.start_epoch
        self.best_epoch = self.start_epoch
        self.stop_fn_flag = False

    def reset(self) -> None:
        self.env_step = 0
        self.start_time = time.time()
        self.train_collector.reset_stat()
        if self.test_collector is not None:
            assert self.episode_per_test is not None
            self.test_collector.reset_stat()
        self.epoch = self.start_epoch
        self.stop_fn_flag = False

    def __iter__(self):
        self.reset()
        return self

    def __next__(self) -> Tuple[int, Dict, Dict]:
        self.epoch += 1
        if self.epoch > self.max_epoch or self.stop_fn_flag:
            raise StopIteration
        self.policy.train()
        with _Bar(self.step_per_epoch, f"Epoch #{self.epoch}", self.show_progress) as t:
            while t.n < t.total:
                stats_train = self.train_step()
                t.update(stats_train["n/st"])
                self.policy_update_fn(stats_train)
                t.set_postfix(cost=stats_train["cost"], rew=stats_train["rew"], length=stats_train["len"])
                self.logger.write_without_reset(self.env_step)
        if self.test_collector is not None:
            self.test_step()
        update_info = self.gather_update_info()
        self.logger.store(tab="update", **update_info)
        if self.epoch % self.save_model_interval == 0:
            self.logger.save_checkpoint()
        if self.perf_is_better(test=True):
            self.logger.save_checkpoint(suffix="best")
        if self.stop_fn and self.stop_fn(self.best_perf_rew, self.best_perf_cost):
            self.stop_fn_flag = True
            self.logger.print("Early stop due to the stop_fn met.", "red")
        epoch_stats = self.logger.stats_mean
        self.logger.write(self.env_step, display=self.verbose)
        update_info.update({"best_reward": self.best_perf_rew, "best_cost": self.best_perf_cost})
        return self.epoch, epoch_stats, update_info

    def perf_is_better(self, test: bool = True) -> bool:
        mode = "test" if test and self.test_collector is not None else "train"
        rew = self.logger.get_mean(mode + "/reward")
        cost = self.logger.get_mean(mode + "/cost")
        feasible_before = self.best_perf_cost <= self.cost_limit
        if not feasible_before:
            better = cost <= self.cost_limit or rew > self.best_perf_rew
        else:
            better = cost <= self.cost_limit and rew > self.best_perf_rew
        if better:
            self.best_perf_cost, self.best_perf_rew = cost, rew
        return better

    def test_step(self) -> Dict[str, Any]:
        assert self.episode_per_test is not None and self.test_collector is not None
        self.test_collector.reset_env()
        self.test_collector.reset_buffer()
        self.policy.eval()
        stats_test = self.test_collector.collect(n_episode=self.episode_per_test)
        self.logger.store(**{"test/reward": stats_test["rew"], "test/cost": stats_test["cost"],
                             "test/length": int(stats_test["len"])})
        return stats_test

    def train_step(self) -> Dict[str, Any]:
        assert self.episode_per_test is not None
        stats_train = self.train_collector.collect(self.episode_per_collect)
        self.env_step += int(stats_train["n/st"])
        self.cum_cost += stats_train["total_cost"]
        self.cum_episode += int(stats_train["n/ep"])
        self.logger.store(**{"update/episode": self.cum_episode, "update/cum_cost": self.cum_cost,
                             "train/reward": stats_train["rew"], "train/cost": stats_train["cost"],
                             "train/length": int(stats_train["len"])})
        return stats_train

    @abstractmethod
    def policy_update_fn(self, result: Dict[str, Any]) -> None:
        ...

    def run(self) -> Dict[str, Union[float, str]]:
        deque(self, maxlen=0)
        return self.gather_update_info()

    def gather_update_info(self) -> Dict[str, Any]:
        duration = max(0, time.time() - self.start_time)
        model_time = max(0, duration - self.train_collector.collect_time)
        result = {"duration": duration}
        if self.test_collector is not None:
            collect_test = self.test_collector.collect_time
            model_time = max(0, model_time - collect_test)
            result.update({"test_time": collect_test,
                           "test_speed": self.test_collector.collect_step / collect_test,
                           "duration": duration})
            train_speed = self.train_collector.collect_step / (duration - collect_test)
        else:
            train_speed = self.train_collector.collect_step / duration
        result.update({"train_collector_time": self.train_collector.collect_time,
                       "train_model_time": model_time, "train_speed": train_speed,
                       "remaining_epoch": self.max_epoch - self.epoch})
        return result
