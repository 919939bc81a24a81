"""Host-side driver of the training loop.  It keeps the protocol of the reference trainer
(/root/reference/fsrl/trainer/base_trainer.py:177-356) -- ``for epoch, stats, info in trainer`` --
and its throughput metric ``train_speed`` = collected env steps / (wall time - evaluation time)
(:345-347), which is the metric BASELINE.json names; everything it schedules runs on the GPU.

An epoch is a sequence of *cycles*; one cycle = collect -> pre_update_fn -> learn -> post_update_fn.
Cycles repeat until the epoch's step quota is spent, then the policy is evaluated, checkpoints are
written and the logger flushes.  Under data parallelism (``policy._dp`` set by fsrl_b200.parallel.attach)
the quota bookkeeping uses step counts agreed between the ranks, so every rank runs the same number of
cycles and issues the same number of gradient exchanges.
"""
from __future__ import annotations

import time
from abc import ABC, abstractmethod
from typing import Any, Callable, Dict, Optional, Tuple, Union

import numpy as np

from ..utils.logger import BaseLogger, DummyLogger


class _Quota:
    """Step budget of one epoch (stands in for the progress bar the reference updates)."""

    def __init__(self, total: int, label: str, echo: bool):
        self.total, self.n, self.label, self.echo = total, 0, label, echo
        self.note: Dict[str, Any] = {}

    def spent(self) -> bool:
        return self.n >= self.total

    def close(self) -> None:
        if self.echo:
            print("%s: %d/%d %s" % (self.label, self.n, self.total, self.note))


class BaseTrainer(ABC):
    def __init__(self, learning_type: str, policy, train_collector, test_collector=None,
                 max_epoch: int = 100, batch_size: int = 512, cost_limit: float = np.inf,
                 step_per_epoch: Optional[int] = None, repeat_per_collect: Optional[int] = None,
                 update_per_step: Union[int, float] = 1, save_model_interval: int = 1,
                 episode_per_test: Optional[int] = None, episode_per_collect: int = 1,
                 stop_fn: Optional[Callable[[float, float], bool]] = None,
                 resume_from_log: bool = False, logger: BaseLogger = DummyLogger(),
                 verbose: bool = True, show_progress: bool = True):
        # what is trained and how it is fed
        self.learning_type = learning_type
        self.policy = policy
        self.train_collector = train_collector
        self.test_collector = test_collector
        self.logger = logger
        # schedule
        self.max_epoch, self.step_per_epoch = max_epoch, step_per_epoch
        self.episode_per_collect, self.episode_per_test = episode_per_collect, episode_per_test
        self.repeat_per_collect, self.update_per_step = repeat_per_collect, update_per_step
        self.batch_size = batch_size
        self.save_model_interval = save_model_interval
        self.stop_fn = stop_fn
        self.resume_from_log = resume_from_log
        self.verbose, self.show_progress = verbose, show_progress
        # model selection: best (reward, cost) pair seen at evaluation time
        self.cost_limit = cost_limit
        self.best_perf_rew, self.best_perf_cost = -np.inf, np.inf
        # counters
        self.start_epoch = 0
        self.epoch = self.best_epoch = self.start_epoch
        self.env_step = 0
        self.cum_cost = 0
        self.cum_episode = 0
        self.stop_fn_flag = False
        self.start_time = time.time()

    # ---- iterator protocol ---------------------------------------------------------------------------
    def reset(self) -> None:
        """Start of a run: zero the step counter, the clock and the collectors' statistics."""
        self.env_step = 0
        self.epoch = self.start_epoch
        self.stop_fn_flag = False
        self.start_time = time.time()
        self.train_collector.reset_stat()
        if self.test_collector is not None:
            assert self.episode_per_test is not None
            self.test_collector.reset_stat()

    def __iter__(self):
        self.reset()
        return self

    def __next__(self) -> Tuple[int, Dict, Dict]:
        if self.stop_fn_flag or self.epoch + 1 > self.max_epoch:
            self.epoch += 1
            raise StopIteration
        self.epoch += 1
        self.policy.train()
        self._train_until_quota()
        return self._close_epoch()

    def _agreed_steps(self, stats_train: Dict[str, Any]) -> int:
        """Env steps this cycle advances the epoch quota by.  A single process uses its own count; data-parallel
        ranks use the maximum over ranks (one small all-reduce), so that no rank leaves the cycle loop -- and
        stops joining gradient exchanges -- before the others."""
        dp = getattr(self.policy, "_dp", None)
        local = int(stats_train.get("local", stats_train)["n/st"])
        if dp is None or dp.world <= 1:
            return int(stats_train["n/st"])
        return int(dp.all_max([local])[0])

    def _train_until_quota(self) -> None:
        quota = _Quota(self.step_per_epoch, "Epoch #%d" % self.epoch, self.show_progress)
        while not quota.spent():
            stats_train = self.train_step()
            self._cycle_steps = self._agreed_steps(stats_train)     # off-policy: also sizes the number of gradient steps
            quota.n += self._cycle_steps
            self.policy_update_fn(stats_train)
            quota.note = dict(cost=stats_train["cost"], rew=stats_train["rew"], length=stats_train["len"])
            self.logger.write_without_reset(self.env_step)
        quota.close()

    def _close_epoch(self) -> Tuple[int, Dict, Dict]:
        if self.test_collector is not None:
            self.test_step()
        info = self.gather_update_info()
        self.logger.store(tab="update", **info)
        if self.epoch % self.save_model_interval == 0:
            self.logger.save_checkpoint()
        if self.perf_is_better(test=True):
            self.logger.save_checkpoint(suffix="best")
        if self.stop_fn is not None and self.stop_fn(self.best_perf_rew, self.best_perf_cost):
            self.stop_fn_flag = True
            self.logger.print("Early stop due to the stop_fn met.", "red")
        epoch_stats = self.logger.stats_mean
        self.logger.write(self.env_step, display=self.verbose)
        info.update(best_reward=self.best_perf_rew, best_cost=self.best_perf_cost)
        return self.epoch, epoch_stats, info

    def run(self) -> Dict[str, Union[float, str]]:
        for _ in self:
            pass
        return self.gather_update_info()

    # ---- one collect / one evaluation ------------------------------------------------------------------
    def train_step(self) -> Dict[str, Any]:
        assert self.episode_per_test is not None
        stats = self.train_collector.collect(self.episode_per_collect)
        self.env_step += int(stats["n/st"])
        self.cum_cost += stats["total_cost"]
        self.cum_episode += int(stats["n/ep"])
        record = {"update/episode": self.cum_episode, "update/cum_cost": self.cum_cost,
                  "train/reward": stats["rew"], "train/cost": stats["cost"], "train/length": int(stats["len"])}
        self.logger.store(**record)
        return stats

    def test_step(self) -> Dict[str, Any]:
        assert self.episode_per_test is not None and self.test_collector is not None
        evaluator = self.test_collector
        evaluator.reset_env()
        evaluator.reset_buffer()
        self.policy.eval()
        stats = evaluator.collect(n_episode=self.episode_per_test)
        record = {"test/reward": stats["rew"], "test/cost": stats["cost"], "test/length": int(stats["len"])}
        self.logger.store(**record)
        return stats

    @abstractmethod
    def policy_update_fn(self, result: Dict[str, Any]) -> None:
        ...

    # ---- model selection and throughput -------------------------------------------------------------------
    def perf_is_better(self, test: bool = True) -> bool:
        """Constrained comparison: an infeasible incumbent is replaced by anything feasible or more rewarding; a
        feasible incumbent only by a feasible candidate with more reward."""
        source = "test" if (test and self.test_collector is not None) else "train"
        reward = self.logger.get_mean(source + "/reward")
        cost = self.logger.get_mean(source + "/cost")
        candidate_ok = cost <= self.cost_limit
        more_reward = reward > self.best_perf_rew
        if self.best_perf_cost <= self.cost_limit:
            accept = candidate_ok and more_reward
        else:
            accept = candidate_ok or more_reward
        if accept:
            self.best_perf_rew, self.best_perf_cost = reward, cost
        return accept

    def gather_update_info(self) -> Dict[str, Any]:
        """Wall-clock split of the run so far.  ``train_speed`` is the headline metric (env steps per second of
        collect + update time, evaluation excluded)."""
        elapsed = max(0, time.time() - self.start_time)
        collecting = self.train_collector.collect_time
        learning = max(0, elapsed - collecting)
        info: Dict[str, Any] = {"duration": elapsed}
        training_wall = elapsed
        if self.test_collector is not None:
            evaluating = self.test_collector.collect_time
            learning = max(0, learning - evaluating)
            training_wall = elapsed - evaluating
            info["test_time"] = evaluating
            info["test_speed"] = self.test_collector.collect_step / evaluating
            info["duration"] = elapsed
        info["train_collector_time"] = collecting
        info["train_model_time"] = learning
        info["train_speed"] = self.train_collector.collect_step / training_wall
        info["remaining_epoch"] = self.max_epoch - self.epoch
        return info
