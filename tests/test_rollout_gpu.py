"""Rollout collection parity: the fused CUDA step kernel (fsrl_rollout_steps through the
C-ABI, via FastCollector) against the oracle restatement of FastCollector.collect
(fast_collector.py:192-408) over the CPU env twin, with identical weights and the same
Philox noise stream.

Tolerances.  The env arithmetic is IEEE-exact on both sides (bit-exact given equal actions);
the actor MLP sums in a different order than torch-CPU and tanhf/expf differ by a few ulp, so
actions agree to ~1e-6 and trajectories to ~1e-4 over an episode.  Integer outcomes (episode
counts, step counts, done flags, buffer pointers) must match exactly; cost flags may flip
only where |x| sits within float noise of the threshold (we allow <= 0.2 % of the steps)."""
import numpy as np
import pytest
import torch

from helpers import buffer_to_numpy, build_ppo, oracle_nets

pytestmark = pytest.mark.gpu

TASKS = ["SafetyCarCircle-v0", "SafetyCarRun-v0", "SafetyBallCircle-v0", "SafetyBallRun-v0",
         "SafetyAntCircle-v0", "SafetyPointGoal1Gymnasium-v0"]


def _oracle_collect(policy, venv, hidden, n_episode, mode="train", collects=1):
    from oracle import collector as ocol
    from oracle.envs import OracleVecEnv
    actor, _ = oracle_nets(policy, hidden)
    oenv = OracleVecEnv(venv.kind, venv.env_num, venv.seed_value)
    oenv.reset()
    obuf = ocol.OracleBuffer(venv.env_num * venv.max_episode_steps * 1, venv.env_num, venv.D, venv.A)
    ctr = np.zeros(venv.env_num, np.uint32)
    out = []
    for _ in range(collects):
        obuf.reset()
        out.append(ocol.collect(oenv, actor, n_episode, policy._act_seed, ctr, obuf, mode=mode))
    return out, obuf, oenv


@pytest.mark.parametrize("task", TASKS)
def test_env_reset_matches_twin_bitwise(task):
    policy, venv, buf, col = build_ppo(task, n_env=33)
    from oracle.envs import OracleVecEnv
    oenv = OracleVecEnv(venv.kind, 33, venv.seed_value)
    obs = oenv.reset()
    assert np.array_equal(venv.obs_cur.cpu().numpy(), obs)
    assert np.array_equal(venv.env_state.cpu().numpy(), oenv.st)


@pytest.mark.parametrize("task", TASKS)
def test_single_collect_matches_oracle(task):
    hidden = (64, 64)
    E = 6
    policy, venv, buf, col = build_ppo(task, hidden=hidden, n_env=E)
    stats = col.collect(n_episode=E)
    (ostats,), obuf, _ = _oracle_collect(policy, venv, hidden, E)
    for k in ("n/ep", "n/st"):
        assert stats[k] == ostats[k]
    assert stats["truncated"] == ostats["truncated"] and stats["terminated"] == ostats["terminated"]
    b = buffer_to_numpy(buf)
    assert np.array_equal(b["len"], obuf.len) and np.array_equal(b["ptr"], obuf.ptr)
    assert np.array_equal(b["terminated"], obuf.terminated) and np.array_equal(b["truncated"], obuf.truncated)
    # first step: identical reset obs -> actions equal to float noise
    T = venv.max_episode_steps
    first = np.arange(E) * obuf.cap
    assert np.array_equal(b["obs"][first], obuf.obs[first])
    np.testing.assert_allclose(b["act"][first], obuf.act[first], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(b["logp"][first], obuf.logp[first], rtol=1e-4, atol=1e-4)
    # whole episode
    if "PointGoal" in task:
        # 1000-step episodes with discontinuous pseudo-lidar bins and goal resampling: float noise in
        # the actions can flip a bin; compare the first 100 steps tightly, the rest statistically
        head = (np.arange(E)[:, None] * obuf.cap + np.arange(100)[None]).ravel()
        np.testing.assert_allclose(b["act"][head], obuf.act[head], rtol=0, atol=5e-3)
        assert (np.abs(b["obs"] - obuf.obs) > 5e-3).mean() < 0.02
    else:
        np.testing.assert_allclose(b["obs"], obuf.obs, rtol=0, atol=5e-3)
        np.testing.assert_allclose(b["act"], obuf.act, rtol=0, atol=5e-3)
        np.testing.assert_allclose(b["rew"], obuf.rew, rtol=0, atol=5e-3)
    assert (b["cost"] != obuf.cost).mean() <= 0.002
    assert abs(stats["rew"] - ostats["rew"]) <= 1e-2 * max(1.0, abs(ostats["rew"]))
    assert abs(stats["cost"] - ostats["cost"]) <= 1.0


def test_env_step_bitexact_given_same_actions():
    """Replay the DEVICE actions through the CPU twin: every stored obs/rew/cost must be
    bit-identical (the env model uses only IEEE-exact ops on both sides)."""
    from oracle.envs import OracleVecEnv
    for task in TASKS:
        E = 5
        policy, venv, buf, col = build_ppo(task, n_env=E)
        col.collect(n_episode=E)
        b = buffer_to_numpy(buf)
        oenv = OracleVecEnv(venv.kind, E, venv.seed_value)
        obs = oenv.reset()
        T = venv.max_episode_steps
        for t in range(T):
            p = np.arange(E) * buf.cap + t
            assert np.array_equal(b["obs"][p], obs), (task, t)
            a = np.clip(b["act"][p], -1, 1).astype(np.float32)
            # map_action scaling with low=-1, high=1 (not the identity in floating point)
            a = (np.float32(-1) + (np.float32(2) * (a + np.float32(1))) / np.float32(2)).astype(np.float32)
            obs, rew, cost, term, trunc = oenv.step(a)
            assert np.array_equal(b["rew"][p], rew) and np.array_equal(b["cost"][p], cost), (task, t)
            assert np.array_equal(b["obs_next"][p], obs)
            assert np.array_equal(b["truncated"][p], trunc)


@pytest.mark.parametrize("n_env,n_episode", [(4, 20), (3, 7), (5, 5), (8, 3)])
def test_episode_count_semantics(n_env, n_episode):
    """Shrinking ready set / surplus rule (fast_collector.py:233-236,355-363)."""
    hidden = (64, 64)
    policy, venv, buf, col = build_ppo("SafetyBallRun-v0", hidden=hidden, n_env=n_env,
                                       buffer_size=n_env * 100 * 8)
    stats = col.collect(n_episode=n_episode)
    from oracle import collector as ocol
    from oracle.envs import OracleVecEnv
    actor, _ = oracle_nets(policy, hidden)
    oenv = OracleVecEnv(venv.kind, n_env, venv.seed_value); oenv.reset()
    obuf = ocol.OracleBuffer(n_env * 100 * 8, n_env, venv.D, venv.A)
    ostats = ocol.collect(oenv, actor, n_episode, policy._act_seed, np.zeros(n_env, np.uint32), obuf)
    assert stats["n/ep"] == ostats["n/ep"] == n_episode
    assert stats["n/st"] == ostats["n/st"]
    b = buffer_to_numpy(buf)
    assert np.array_equal(b["len"], obuf.len)
    assert np.array_equal(b["truncated"], obuf.truncated)
    np.testing.assert_allclose(stats["len"], ostats["len"])
    # after the collect all envs were reset: same episode counters, same fresh obs
    assert np.array_equal(venv.ep_idx.cpu().numpy().astype(np.uint32), oenv.ep_idx)
    assert np.array_equal(venv.obs_cur.cpu().numpy(), oenv.observe())


def test_eval_mode_is_deterministic_and_random_mode_runs():
    policy, venv, buf, col = build_ppo("SafetyCarCircle-v0", n_env=4)
    policy.eval()
    s1 = col.collect(n_episode=4)
    a1 = buf.act.clone()
    venv.ep_idx.zero_(); col.reset_env(); col.reset_buffer()
    s2 = col.collect(n_episode=4)
    assert torch.equal(a1, buf.act) and s1["rew"] == s2["rew"]
    policy.train()
    s3 = col.collect(n_episode=4, random=True)
    assert s3["n/ep"] == 4 and float(buf.act.abs().max()) <= 1.0


def test_second_collect_continues_rng_streams():
    hidden = (64, 64)
    policy, venv, buf, col = build_ppo("SafetyBallCircle-v0", hidden=hidden, n_env=4)
    col.collect(n_episode=4); col.reset_buffer()
    stats = col.collect(n_episode=4)
    outs, obuf, _ = _oracle_collect(policy, venv, hidden, 4, collects=2)
    b = buffer_to_numpy(buf)
    first = np.arange(4) * obuf.cap
    assert np.array_equal(b["obs"][first], obuf.obs[first])
    np.testing.assert_allclose(b["act"][first], obuf.act[first], rtol=2e-5, atol=2e-6)
    assert stats["n/st"] == outs[1]["n/st"]
