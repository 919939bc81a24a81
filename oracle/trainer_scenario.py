"""Deterministic fake collector / policy used to compare trainer implementations (test infrastructure).

``run(OnTrainer, OffTrainer, LoggerCls)`` drives an on-policy and an off-policy trainer for a few
epochs with scripted collect statistics and records (a) every call the trainer makes on the policy and
the collectors, in order, with its arguments, and (b) what it returns / logs, minus wall-clock values.
oracle/make_golden_policies.py runs it with the REFERENCE's trainers and logger
(fsrl.trainer.*, fsrl.utils.BaseLogger) and stores the record as tests/golden/trainer_golden.json;
tests/test_oracle_golden.py runs it with fsrl_b200's and compares.
"""
from __future__ import annotations

import numpy as np

TIME_KEYS = ("duration", "train_collector_time", "train_model_time", "train_speed", "test_time", "test_speed")


class FakeCollector:
    def __init__(self, name, trace, seed):
        self.name, self.trace = name, trace
        self.rng = np.random.default_rng(seed)
        self.buffer = "buffer:" + name
        self.collect_step = self.collect_episode = 0
        self.collect_time = 0.0

    def reset_stat(self):
        self.trace.append([self.name, "reset_stat"])
        self.collect_step = self.collect_episode = 0
        self.collect_time = 0.0

    def reset_env(self, *a, **k):
        self.trace.append([self.name, "reset_env"])

    def reset_buffer(self, keep_statistics=False):
        self.trace.append([self.name, "reset_buffer", bool(keep_statistics)])

    def collect(self, n_episode=None, **kw):
        self.trace.append([self.name, "collect", n_episode, sorted(kw)])
        lens = self.rng.integers(20, 60, size=n_episode)
        n_st = int(lens.sum())
        cost = float(np.round(self.rng.uniform(0, 30), 3))
        rew = float(np.round(self.rng.uniform(-5, 40), 3))
        self.collect_step += n_st
        self.collect_episode += n_episode
        self.collect_time += 0.01
        return {"n/ep": n_episode, "n/st": n_st, "rew": rew, "len": float(lens.mean()), "total_cost": cost * n_episode,
                "cost": cost, "truncated": 1.0, "terminated": 0.0}


class FakePolicy:
    def __init__(self, trace):
        self.trace = trace

    def train(self, mode=True):
        self.trace.append(["policy", "train"])
        return self

    def eval(self):
        self.trace.append(["policy", "eval"])
        return self

    def pre_update_fn(self, **kw):
        self.trace.append(["policy", "pre_update_fn", {k: (v if isinstance(v, (int, float, str)) else "<obj>")
                                                       for k, v in sorted(kw.items()) if k != "stats_train"},
                           kw["stats_train"]["n/st"]])

    def update(self, sample_size, buffer, **kw):
        self.trace.append(["policy", "update", sample_size, buffer, {k: v for k, v in sorted(kw.items())}])

    def post_update_fn(self, **kw):
        self.trace.append(["policy", "post_update_fn", sorted(kw)])


def _clean(d):
    out = {}
    for k, v in d.items():
        if any(k.endswith(t) for t in TIME_KEYS):
            continue
        out[k] = float(v) if isinstance(v, (int, float, np.floating, np.integer)) else v
    return out


def run(OnTrainer, OffTrainer, LoggerCls):
    record = {}
    for kind in ("onpolicy", "offpolicy"):
        trace = []
        policy = FakePolicy(trace)
        train_c, test_c = FakeCollector("train", trace, 1), FakeCollector("test", trace, 2)
        logger = LoggerCls()
        stops = []

        def stop_fn(best_rew, best_cost):
            stops.append([float(best_rew), float(best_cost)])
            return len(stops) >= 3                         # ends the run after the third epoch

        common = dict(max_epoch=5, batch_size=64, cost_limit=12.0, step_per_epoch=400, episode_per_collect=4,
                      episode_per_test=2, save_model_interval=2, stop_fn=stop_fn, logger=logger, verbose=False,
                      show_progress=False)
        if kind == "onpolicy":
            trainer = OnTrainer(policy, train_c, test_c, repeat_per_collect=3, **common)
        else:
            trainer = OffTrainer(policy, train_c, test_c, update_per_step=0.05, **common)
        epochs = []
        for epoch, stats, info in trainer:
            epochs.append({"epoch": int(epoch), "stats": _clean(stats), "info": _clean(info)})
        record[kind] = {"trace": trace, "epochs": epochs, "stops": stops,
                        "env_step": int(trainer.env_step), "cum_episode": int(trainer.cum_episode),
                        "cum_cost": float(trainer.cum_cost)}
    return record


def logger_files(LoggerCls, log_dir):
    """Drive a file-backed logger through config saving, two writes and a checkpoint; return the text of
    the files it produced (config.yaml, progress.txt) and the checkpoint file names."""
    import os
    import torch
    lg = LoggerCls(log_dir, log_txt=True, name="run")
    cfg = {"task": "SafetyCarCircle-v0", "hidden_sizes": (128, 128), "lr": 5e-4, "lagrangian_pid": (0.05, 0.0005, 0.1),
           "cost_limit": 10, "unbounded": False, "group": None}
    lg.save_config(cfg, verbose=False)
    lg.setup_checkpoint_fn(lambda: {"model": {"w": torch.arange(3.0)}})
    for step, vals in ((1000, [(1.0, 12.0), (3.0, 8.0)]), (2000, [(5.0, 4.0)])):
        for rew, cost in vals:
            lg.store(tab="train", reward=rew, cost=cost)
        lg.store(tab="update", gradient_steps=step // 10)
        lg.store(total=0.25 * step, tab="loss")
        lg.write(step, display=False)
    lg.save_checkpoint()
    lg.save_checkpoint(suffix=7)
    lg.save_checkpoint(suffix="best")
    if getattr(lg, "output_file", None) is not None:
        lg.output_file.flush()
    run = os.path.join(log_dir, "run")
    return {"cfg_after": {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()},
            "config.yaml": open(os.path.join(run, "config.yaml")).read(),
            "progress.txt": open(os.path.join(run, "progress.txt")).read(),
            "checkpoints": sorted(os.listdir(os.path.join(run, "checkpoint")))}


class TerminatingEnv:
    """The numpy env twin plus a scripted termination rule (our env models only ever truncate at T): env e
    terminates at step t of its k-th episode when (7 e + 13 t + 5 k) % period == 0, so that episodes end at
    different times in different envs -- what the collector's surplus-env bookkeeping has to cope with."""

    def __init__(self, kind, n_env, seed, period=41):
        from oracle.envs import OracleVecEnv
        self.inner = OracleVecEnv(kind, n_env, seed)
        self.period = period
        self.E, self.D, self.A = self.inner.E, self.inner.D, self.inner.A

    def reset(self, ids=None):
        return self.inner.reset(ids)

    def observe(self, ids=None):
        return self.inner.observe(ids)

    def step(self, act, ids):
        ids = np.asarray(ids)
        k = self.inner.ep_idx[ids].astype(np.int64)
        obs_next, rew, cost, term, trunc = self.inner.step(act, ids)
        t = self.inner.t[ids].astype(np.int64)
        term = term | ((7 * ids.astype(np.int64) + 13 * t + 5 * k) % self.period == 0)
        return obs_next, rew, cost, term, trunc
