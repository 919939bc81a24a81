"""Device-resident vector environments behind the vector-env protocol the reference's
collector consumes (``len(env)``, ``env.action_space``, ``reset``/``step`` --
fsrl/data/fast_collector.py:134,172,286; tianshou BaseVectorEnv).

The dynamics are the analytic models of csrc/envs.cuh (bullet_safety_gym / safety_gymnasium
are absent and irreproducible).  All state lives in HBM as SoA tensors; the
fused rollout kernel (csrc/rollout.cu) steps every env without host involvement.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from .spaces import Box

# task id -> (kind, D, A, S, T).  Registry names follow the reference's examples
# (examples/mlp/train_ppol_agent.py:28-40); BASELINE.json's "SafetyPointGoal1-v0" is an alias
# of the registry name SafetyPointGoal1Gymnasium-v0.
KINDS = {
    "SafetyCarCircle-v0": 0, "SafetyCarRun-v0": 1, "SafetyBallCircle-v0": 2, "SafetyBallRun-v0": 3,
    "SafetyAntCircle-v0": 4, "SafetyPointGoal1Gymnasium-v0": 5, "SafetyPointGoal1-v0": 5,
}


def env_dims(kind: int):
    D, A, S, T = (ctypes.c_int() for _ in range(4))
    _lib.check(_lib.lib.fsrl_env_dims(kind, ctypes.byref(D), ctypes.byref(A), ctypes.byref(S), ctypes.byref(T)))
    return D.value, A.value, S.value, T.value


class _Spec:
    def __init__(self, id, max_episode_steps):
        self.id, self.max_episode_steps = id, max_episode_steps


class DeviceEnv:
    """What ``gym.make(task)`` returns: the spaces + horizon of one env (a descriptor: stepping
    happens only inside a :class:`DeviceVectorEnv`)."""

    def __init__(self, task: str):
        if task not in KINDS:
            raise KeyError(f"unknown task {task!r}; available: {sorted(KINDS)}")
        self.task = task
        self.kind = KINDS[task]
        D, A, S, T = env_dims(self.kind)
        self.observation_space = Box(-np.inf, np.inf, (D,), np.float32)
        self.action_space = Box(-1.0, 1.0, (A,), np.float32)
        self.spec = _Spec(task, T)
        self.state_dim = S

    def close(self):
        pass


def make(task: str, **_) -> DeviceEnv:
    return DeviceEnv(task)


class DeviceVectorEnv:
    """E independent envs of one task, resident on one GPU."""

    def __init__(self, task: str, env_num: int, device="cuda", seed: int = 0):
        proto = DeviceEnv(task)
        self.task, self.kind = task, proto.kind
        self.env_num = int(env_num)
        self.device = torch.device(device)
        self.observation_space = proto.observation_space
        self.action_space = proto.action_space
        self.max_episode_steps = proto.spec.max_episode_steps
        self.spec = proto.spec
        self.seed_value = int(seed) & 0xFFFFFFFF
        D, A, S, T = env_dims(self.kind)
        self.D, self.A, self.S = D, A, S
        E, dev = self.env_num, self.device
        self.env_state = torch.zeros((S, E), dtype=torch.float32, device=dev)
        self.obs_cur = torch.zeros((E, D), dtype=torch.float32, device=dev)
        self.env_t = torch.zeros(E, dtype=torch.int32, device=dev)
        self.ep_idx = torch.zeros(E, dtype=torch.int32, device=dev)
        self.act_ctr = torch.zeros(E, dtype=torch.int32, device=dev)
        self.active = torch.zeros(E, dtype=torch.uint8, device=dev)
        self.done_now = torch.zeros(E, dtype=torch.uint8, device=dev)
        self.ep_rew = torch.zeros(E, dtype=torch.float64, device=dev)
        self.ep_len = torch.zeros(E, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(ctypes.sizeof(_lib.CollectStats), dtype=torch.uint8, device=dev)
        self._stats_host = torch.zeros(ctypes.sizeof(_lib.CollectStats), dtype=torch.uint8).pin_memory() \
            if torch.cuda.is_available() else torch.zeros(ctypes.sizeof(_lib.CollectStats), dtype=torch.uint8)

    def __len__(self):
        return self.env_num

    def seed(self, seed=None):
        if seed is not None:
            self.seed_value = int(seed) & 0xFFFFFFFF
        return [self.seed_value] * self.env_num

    # ---- descriptor shared by every rollout entry point ----------------------------------------
    def fill(self, r: "_lib.Rollout") -> None:
        r.kind, r.E, r.max_steps = self.kind, self.env_num, self.max_episode_steps
        r.seed_env = self.seed_value
        r.env_state, r.obs_cur = self.env_state.data_ptr(), self.obs_cur.data_ptr()
        r.env_t, r.ep_idx, r.act_ctr = self.env_t.data_ptr(), self.ep_idx.data_ptr(), self.act_ctr.data_ptr()
        r.active, r.done_now = self.active.data_ptr(), self.done_now.data_ptr()
        r.ep_rew, r.ep_len = self.ep_rew.data_ptr(), self.ep_len.data_ptr()
        r.stats = self.stats.data_ptr()
        low, high = self.action_space.low, self.action_space.high
        for j in range(self.A):
            r.act_low[j], r.act_high[j] = float(low[j]), float(high[j])

    def reset(self, ids=None, **kwargs):
        """Start a fresh episode in every env (partial resets happen inside the rollout
        kernel; ``ids`` other than None/all is not part of the device protocol)."""
        if ids is not None and len(ids) != self.env_num:
            raise NotImplementedError("DeviceVectorEnv resets individual envs on the device only")
        r = _lib.Rollout()
        self.fill(r)
        r.mode = _lib.MODE_RANDOM
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.fsrl_env_reset_all(ctypes.byref(r), torch.cuda.current_stream().cuda_stream))
        return self.obs_cur, [{} for _ in range(self.env_num)]

    def read_stats(self) -> "_lib.CollectStats":
        self._stats_host.copy_(self.stats, non_blocking=False)
        return _lib.CollectStats.from_buffer_copy(self._stats_host.numpy().tobytes())

    def step(self, action, id=None):
        raise NotImplementedError(
            "DeviceVectorEnv is stepped by the fused rollout kernel (FastCollector.collect); "
            "a host-side step() would round-trip every action through PCIe")

    def render(self, **kwargs):
        return None

    def close(self):
        pass
