"""``FastCollector`` with the reference's constructor, ``collect`` contract and result keys
(/root/reference/fsrl/data/fast_collector.py:47-69,192-408), running every vector step as
one fused CUDA launch (csrc/rollout.cu) against a :class:`DeviceVectorEnv`.

Episode-count semantics are the reference's: the ready set is the first
``min(env_num, n_episode)`` envs (:235-236), finished envs are reset and keep going until the
remaining episode budget is smaller than the ready set, at which point the lowest-index
finished envs are retired first (:357-363); every collect ends with a reset of all envs
(:375-388).  When ``n_episode <= env_num`` each ready env runs exactly one episode and the
bookkeeping is done inline by the step kernel; otherwise a one-CTA resolve kernel applies
the ordered surplus rule after each step.
"""
from __future__ import annotations

import ctypes
import time
from typing import Any, Callable, Dict, Optional

import torch

from .. import _lib
from ..envs import DeviceVectorEnv
from .batch import Batch
from .buffer import DeviceVectorReplayBuffer


class FastCollector(object):
    def __init__(self, policy, env: DeviceVectorEnv, buffer: Optional[DeviceVectorReplayBuffer] = None,
                 preprocess_fn: Optional[Callable[..., Batch]] = None,
                 exploration_noise: bool = False) -> None:
        super().__init__()
        if not isinstance(env, DeviceVectorEnv):
            raise TypeError("fsrl_b200.FastCollector steps DeviceVectorEnv instances on the GPU; "
                            f"got {type(env).__name__}")
        if preprocess_fn is not None:
            raise NotImplementedError("preprocess_fn would need a host round trip per step")
        self.env = env
        self.env_num = len(env)
        self.exploration_noise = exploration_noise
        self._store = buffer is not None
        self._assign_buffer(buffer)
        self.policy = policy
        self.preprocess_fn = None
        self._action_space = env.action_space
        self.reset(False)

    def _assign_buffer(self, buffer) -> None:
        if buffer is None:
            # the reference creates VectorReplayBuffer(env_num, env_num) for a buffer-less
            # collector (:72-73) whose content nobody reads (evaluate()); we skip the stores
            self.buffer = None
            return
        assert buffer.buffer_num >= self.env_num                        # :75
        buffer.allocate(self.env.D, self.env.A, self.env.device)
        self.buffer = buffer

    def reset(self, reset_buffer: bool = True, gym_reset_kwargs: Optional[Dict[str, Any]] = None) -> None:
        self.reset_env(gym_reset_kwargs)
        if reset_buffer:
            self.reset_buffer()
        self.reset_stat()

    def reset_stat(self) -> None:
        self.collect_step, self.collect_episode, self.collect_time = 0, 0, 0.0

    def reset_buffer(self, keep_statistics: bool = False) -> None:
        if self.buffer is not None:
            self.buffer.reset(keep_statistics=keep_statistics)

    def reset_env(self, gym_reset_kwargs: Optional[Dict[str, Any]] = None) -> None:
        self.env.reset()

    # ------------------------------------------------------------------------------------------------
    def _descriptor(self, random: bool) -> "_lib.Rollout":
        r = _lib.Rollout()
        self.env.fill(r)
        if self.buffer is not None:
            self.buffer.fill(r)
        if hasattr(self.policy, "fill_rollout"):
            self.policy.fill_rollout(r, exploration_noise=self.exploration_noise)
        elif not random:
            raise TypeError("the policy does not expose fill_rollout(); only random=True "
                            "collection is possible with it")
        else:
            r.actor.H = 64
            r.action_bound = {"": 0, "clip": 1, "tanh": 2}[getattr(self.policy, "action_bound_method", "clip")]
            r.action_scaling = int(getattr(self.policy, "action_scaling", True))
        if random:
            r.mode = _lib.MODE_RANDOM
        return r

    def collect(self, n_episode: int = 1, random: bool = False, render: bool = False,
                no_grad: bool = True, gym_reset_kwargs: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        if n_episode is not None:
            assert n_episode > 0                                        # :234
        else:
            raise TypeError("Please specify n_episode"
                            "in FastCollector.collect().")
        start_time = time.time()
        env = self.env
        r = self._descriptor(random)
        r.inline_done = 1 if n_episode <= self.env_num else 0
        T = env.max_episode_steps
        with torch.cuda.device(env.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib.fsrl_collect_begin(ctypes.byref(r), int(n_episode), stream))
            if r.inline_done:
                # every ready env runs exactly one episode of at most T steps
                _lib.check(_lib.lib.fsrl_rollout_steps(ctypes.byref(r), T, stream))
                st = env.read_stats()
            else:
                chunk = max(1, min(T, 64))
                while True:
                    _lib.check(_lib.lib.fsrl_rollout_steps(ctypes.byref(r), chunk, stream))
                    st = env.read_stats()
                    if st.finished:
                        break
        if not st.finished:
            raise RuntimeError("rollout did not reach n_episode within the step bound "
                               f"(episodes {st.episode_count}/{n_episode})")
        step_count, episode_count = int(st.step_count), int(st.episode_count)
        self.collect_step += step_count
        self.collect_episode += episode_count
        # a collect always ends with fresh resets of every env (:375-388)
        self.reset_env()
        self.collect_time += max(time.time() - start_time, 1e-9)

        if episode_count > 0:
            rew_mean = st.sum_ep_rew / episode_count
            len_mean = st.sum_ep_len / episode_count
        else:
            rew_mean = len_mean = 0
        done_count = st.term_count + st.trunc_count
        return {
            "n/ep": episode_count,
            "n/st": step_count,
            "rew": rew_mean,
            "len": len_mean,
            "total_cost": st.total_cost,
            "cost": st.total_cost / episode_count,
            "truncated": st.trunc_count / done_count,
            "terminated": st.term_count / done_count,
        }
