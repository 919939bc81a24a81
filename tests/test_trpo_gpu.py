"""TRPO-Lagrangian parity: natural-gradient direction (CG on exact Fisher-vector
products), KL-bounded step and backtracking on the device vs the torch-CPU restatement of
trpo_lag.py:182-246.  step_size must match exactly (same number of backtracks); scalars rtol 5e-3
on the first step."""
import numpy as np
import pytest
import torch

from helpers import buffer_to_numpy

pytestmark = pytest.mark.gpu


def test_trpo_learn_matches_oracle():
    from fsrl_b200 import envs
    from fsrl_b200.agent import TRPOLagAgent
    from fsrl_b200.data import FastCollector, VectorReplayBuffer
    from oracle import nets as onets, trpo as otrpo
    hidden = (64, 64)
    task = "SafetyCarCircle-v0"
    env = envs.make(task)
    agent = TRPOLagAgent(env, seed=10, hidden_sizes=hidden, optim_critic_iters=3, target_kl=0.001)
    policy = agent.policy
    venv = envs.DeviceVectorEnv(task, 4, seed=12)
    buf = VectorReplayBuffer(4 * 300, 4)
    col = FastCollector(policy, venv, buf, exploration_noise=True)
    stats = col.collect(n_episode=4)
    policy.lag_optims[0].lagrangian = 0.6
    sd = policy.state_dict()
    D, A = venv.D, venv.A
    actor = onets.load_from_state_dict(onets.GaussActor(D, A, list(hidden)), sd, "actor.")
    critics = [onets.load_from_state_dict(onets.ValueNet(D, list(hidden)), sd, f"critics.{i}.") for i in range(2)]
    idx = buf.sample_indices(0)
    batch = policy.process_fn(None, buf, idx)
    b = buffer_to_numpy(buf)
    sel = idx.cpu().numpy()
    ob = {k: b[k][sel] for k in ("obs", "obs_next", "act", "rew", "cost", "terminated", "truncated")}
    ob = otrpo.process(actor, critics, ob, 0.99, 0.95)
    ob["advs"] = batch.advs.cpu().numpy().copy(); ob["rets"] = batch.rets.cpu().numpy().copy()
    ob["logp_old"] = batch.logp_old.cpu().numpy().copy()
    opt = torch.optim.Adam([p for c in critics for p in c.parameters()], lr=5e-4)
    np.random.seed(4)
    ostats = otrpo.learn(actor, critics, opt, ob, 99999, 2, 0.6, optim_critic_iters=3, delta=0.001)
    np.random.seed(4)
    policy.learn(batch, batch_size=99999, repeat=2)
    st = policy.last_stats
    for k in range(2):
        assert abs(st["loss/step_size"][k] - ostats[k]["loss/step_size"]) <= 5e-3 * abs(ostats[k]["loss/step_size"]) + 1e-9
    for key in ("loss/actor_rew", "loss/actor_safety", "loss/actor_total", "loss/kl", "loss/vf0", "loss/vf1"):
        want = np.array([s[key] for s in ostats]); got = np.array(st[key])
        np.testing.assert_allclose(got[:1], want[:1], rtol=5e-3, atol=2e-5, err_msg=key)
        np.testing.assert_allclose(got, want, rtol=0.15, atol=1e-3, err_msg=key)
