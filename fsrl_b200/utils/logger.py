"""Host-side logging with the reference's logger surface (fsrl/utils/logger/base_logger.py:
``store / write / write_without_reset / save_checkpoint / save_config / stats_mean /
get_mean / print`` and the on-disk formats ``progress.txt`` (TSV), ``config.yaml`` and
``checkpoint/model[_suffix].pt`` = ``{"model": state_dict}``), so the reference's eval and
plotting scripts read our runs.  Pure host code; not on the hot path: the
policies hand it one batch of per-minibatch statistics per update instead of one ``.item()``
per scalar."""
from __future__ import annotations

import atexit
import os
import time
from typing import Callable, Dict, Iterable, Optional, Union

import numpy as np
import torch


class _Welford:
    __slots__ = ("n", "mu", "m2")

    def __init__(self):
        self.n, self.mu, self.m2 = 0, 0.0, 0.0

    def add(self, x) -> None:
        self.n += 1
        d = x - self.mu
        self.mu = self.mu + d / self.n
        self.m2 = self.m2 + d * (x - self.mu)

    @property
    def mean(self):
        return self.mu

    @property
    def std(self):
        return float(np.sqrt(self.m2 / self.n)) if self.n else 0.0


class BaseLogger:
    def __init__(self, log_dir=None, log_txt=True, name=None) -> None:
        self.name = name if name is not None else time.strftime("%Y-%m-%d_exp")
        self.log_dir = os.path.join(log_dir, name) if log_dir is not None else None
        self.log_fname = "progress.txt"
        self.output_file = None
        if self.log_dir:
            os.makedirs(self.log_dir, exist_ok=True)
            if log_txt:
                self.output_file = open(os.path.join(self.log_dir, self.log_fname), "w")
                atexit.register(self.output_file.close)
        self.first_row = True
        self.checkpoint_fn: Optional[Callable] = None
        self.reset_data()

    # ---- data ---------------------------------------------------------------------------------
    def setup_checkpoint_fn(self, checkpoint_fn: Optional[Callable] = None) -> None:
        self.checkpoint_fn = checkpoint_fn

    def reset_data(self) -> None:
        self.log_data: Dict[str, _Welford] = {}

    def store(self, tab: str = None, **kwargs) -> None:
        for k, v in kwargs.items():
            key = k if tab is None else tab + "/" + k
            self.log_data.setdefault(key, _Welford()).add(float(np.mean(v)))

    def store_many(self, tab: Optional[str], key: str, values) -> None:
        """Feed a whole vector of per-minibatch values (one device->host copy per update): the batch's moments are
        merged into the running ones (Chan et al.'s pairwise update) instead of one Python-level add per value --
        9 600 minibatches x 12 keys per cycle cost 35 ms of host time the other way."""
        k = key if tab is None else tab + "/" + key
        w = self.log_data.setdefault(k, _Welford())
        x = np.asarray(values, dtype=np.float64).ravel()
        nb = x.size
        if nb == 0:
            return
        mb = float(x.mean())
        m2b = float(((x - mb) ** 2).sum())
        n = w.n + nb
        d = mb - w.mu
        w.mu = w.mu + d * nb / n
        w.m2 = w.m2 + m2b + d * d * w.n * nb / n
        w.n = n

    @property
    def logger_keys(self) -> Iterable:
        return self.log_data.keys()

    def get_mean(self, key: str) -> float:
        return self.log_data[key].mean if key in self.log_data else 0.0

    def get_std(self, key: str) -> float:
        return self.log_data[key].std if key in self.log_data else 0.0

    def get_mean_list(self, keys: Iterable[str]) -> list:
        return [self.get_mean(k) for k in keys]

    def get_mean_dict(self, keys: Iterable[str]) -> dict:
        return {k: self.get_mean(k) for k in keys}

    @property
    def stats_mean(self) -> dict:
        return self.get_mean_dict(self.logger_keys)

    # ---- sinks -----------------------------------------------------------------------------------
    def write(self, step: int, display: bool = False, display_keys: Iterable[str] = None) -> None:
        if "update/env_step" not in self.log_data:
            self.store(tab="update", env_step=step)
        if self.output_file is not None:
            keys = list(self.logger_keys)
            if self.first_row:
                self.output_file.write("\t".join(["Steps"] + keys) + "\n")
                self.first_row = False
            self.output_file.write("\t".join(map(str, [step] + self.get_mean_list(keys))) + "\n")
            self.output_file.flush()
        if display:
            self.display_tabular(display_keys)
        self.reset_data()

    def write_without_reset(self, *args, **kwargs) -> None:
        pass

    def save_checkpoint(self, suffix: Optional[Union[int, str]] = None) -> None:
        if self.checkpoint_fn and self.log_dir:
            d = os.path.join(self.log_dir, "checkpoint")
            os.makedirs(d, exist_ok=True)
            tag = "" if suffix is None else "_" + (("%d" % suffix) if isinstance(suffix, int) else suffix)
            torch.save(self.checkpoint_fn(), os.path.join(d, "model" + tag + ".pt"))

    def save_config(self, config: dict, verbose=True) -> None:
        """``<log_dir>/config.yaml`` in the reference's wire format (base_logger.py:128-163): the run name is
        written INTO the caller's dict, and the dict is dumped as is (block style, insertion order, tuples as
        ``!!python/tuple``) so that either side's ``load_config_and_model`` reads the other's runs."""
        if self.name is not None:
            config["name"] = self.name
        if verbose:
            print("Saving config:", {k: v for k, v in config.items()})
        if self.log_dir:
            import yaml
            with open(os.path.join(self.log_dir, "config.yaml"), "w") as f:
                try:
                    yaml.dump(config, f, default_flow_style=False, indent=4, sort_keys=False)
                except yaml.representer.RepresenterError:      # an object PyYAML cannot tag: fall back to plain types
                    f.seek(0); f.truncate()
                    yaml.dump(_plain(config), f, default_flow_style=False, indent=4, sort_keys=False)

    def restore_data(self) -> None:
        pass

    def display_tabular(self, display_keys: Iterable[str] = None) -> None:
        keys = sorted(self.logger_keys) if not display_keys else list(display_keys)
        if not keys:
            return
        w = max(15, max(len(k) for k in keys))
        bar = "-" * (w + 22)
        print(bar)
        for k in keys:
            print(f"| {k:>{w}s} | {self.get_mean(k):>15.5g} |")
        print(bar, flush=True)

    def print(self, msg: str, color="green") -> None:
        print(msg)


def _plain(x):
    if isinstance(x, dict):
        return {str(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    if isinstance(x, (np.generic,)):
        return x.item()
    if isinstance(x, (int, float, str, bool)) or x is None:
        return x
    return str(x)


class DummyLogger(BaseLogger):
    """Swallows everything (the reference's default when no logger is configured)."""

    def __init__(self, *args, **kwargs) -> None:
        self.log_dir, self.name, self.output_file, self.checkpoint_fn = None, None, None, None
        self.first_row = True
        self.reset_data()

    def store(self, *args, **kwargs) -> None:
        pass

    def store_many(self, *args, **kwargs) -> None:
        pass

    def write(self, *args, **kwargs) -> None:
        pass

    def save_config(self, *args, **kwargs) -> None:
        pass

    def print(self, *args, **kwargs) -> None:
        pass


class TensorboardLogger(BaseLogger):
    """BaseLogger + a tensorboard event file under ``<log_dir>/tb`` (reference:
    fsrl/utils/logger/tb_logger.py:10-81): every ``write`` also emits one scalar per stored
    key, and ``restore_data`` recovers (epoch, env_step, gradient_step) from the event file
    of an earlier run so that ``agent.learn(resume=True)`` continues the counters."""

    def __init__(self, log_dir: str = None, log_txt: bool = True, name: str = None) -> None:
        super().__init__(log_dir, log_txt, name)
        from torch.utils.tensorboard import SummaryWriter
        self.summary_writer = SummaryWriter(os.path.join(self.log_dir, "tb"))
        self.last_save_step = self.last_log_test_step = -1
        self.last_log_update_step = self.last_log_train_step = -1

    def write(self, step: int, display: bool = True, display_keys: Iterable[str] = None) -> None:
        self.store(tab="update", env_step=step)
        self.write_without_reset(step)
        return super().write(step, display, display_keys)

    def write_without_reset(self, step: int) -> None:
        for key in self.logger_keys:
            self.summary_writer.add_scalar(key, self.get_mean(key), step)
        self.summary_writer.flush()

    def restore_data(self):
        from tensorboard.backend.event_processing import event_accumulator
        acc = event_accumulator.EventAccumulator(self.summary_writer.log_dir)
        acc.Reload()

        def last_step(tag):
            return acc.scalars.Items(tag)[-1].step

        epoch = gradient_step = env_step = 0
        try:
            epoch = last_step("update/episode")
            self.last_save_step = self.last_log_test_step = epoch
            gradient_step = last_step("update/gradient_steps")
            self.last_log_update_step = gradient_step
        except KeyError:
            epoch, gradient_step = 0, 0
        try:
            env_step = last_step("update/env_step")
            self.last_log_train_step = env_step
        except KeyError:
            env_step = 0
        return epoch, env_step, gradient_step


class WandbLogger(BaseLogger):
    """BaseLogger + Weights & Biases (reference: fsrl/utils/logger/wandb_logger.py:9-73): a run is
    opened (or the already active one adopted) at construction and the per-key means are sent
    with every ``write``.  The container has no network: pass ``WANDB_MODE=offline``."""

    def __init__(self, config: dict = {}, project: str = "fsrl", group: str = "test", name: str = None,
                 log_dir: str = "log", log_txt: bool = True) -> None:
        super().__init__(log_dir, log_txt, name)
        import uuid
        import wandb
        self._wandb = wandb
        self.wandb_run = wandb.run if wandb.run else wandb.init(
            project=project, group=group, name=name, id=str(uuid.uuid4()), resume="allow", config=config)

    def write(self, step: int, display: bool = True, display_keys: Iterable[str] = None) -> None:
        self.store(tab="update", env_step=step)
        self.write_without_reset(step)
        return super().write(step, display, display_keys)

    def write_without_reset(self, step: int) -> None:
        self._wandb.log(self.stats_mean, step=step)

    def restore_data(self) -> None:
        """The reference does not restore W&B runs either (wandb_logger.py:72-73)."""
