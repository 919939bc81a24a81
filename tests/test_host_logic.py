"""CPU tests of the host-side logic (no GPU, no compute kernels): replay-buffer index contract vs
the oracle restatement of tianshou's VectorReplayBuffer semantics, the Batch
container, config tables, compat shims, logger on-disk formats."""
import dataclasses
import os
import sys

import numpy as np
import pytest
import torch


def _filled_buffers(E=4, cap=7, seed=0):
    from fsrl_b200.data.buffer import DeviceVectorReplayBuffer
    from oracle.collector import OracleBuffer
    rng = np.random.default_rng(seed)
    buf = DeviceVectorReplayBuffer(E * cap, E, device="cpu")
    buf.allocate(3, 2, "cpu")
    ob = OracleBuffer(E * cap, E, 3, 2)
    steps = rng.integers(2, 2 * cap, E)          # some envs wrap their ring
    for e in range(E):
        for t in range(steps[e]):
            term, trunc = bool(rng.random() < 0.15), bool(rng.random() < 0.1)
            ids = np.array([e])
            ob.add(ids, rng.standard_normal((1, 3)).astype(np.float32), np.zeros((1, 2), np.float32),
                   np.float32([t]), np.float32([0]), np.float32([0]), np.array([term]), np.array([trunc and not term]),
                   np.zeros((1, 3), np.float32))
    buf.terminated.copy_(torch.from_numpy(ob.terminated.astype(np.uint8)))
    buf.truncated.copy_(torch.from_numpy(ob.truncated.astype(np.uint8)))
    buf.rew.copy_(torch.from_numpy(ob.rew))
    buf.ptr.copy_(torch.from_numpy(ob.ptr.astype(np.int32)))
    buf.len.copy_(torch.from_numpy(ob.len.astype(np.int32)))
    return buf, ob


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_buffer_index_contract_matches_oracle(seed):
    from oracle.offpolicy import buffer_next
    buf, ob = _filled_buffers(seed=seed)
    idx = buf.sample_indices(0).numpy()
    assert np.array_equal(idx, ob.sample_all())                       # sub-buffer order, chronological
    assert len(buf) == int(ob.len.sum())
    assert np.array_equal(np.sort(buf.unfinished_index().numpy()), np.sort(ob.unfinished_index()))
    valid = torch.from_numpy(ob.sample_all())
    assert np.array_equal(buf.next(valid).numpy(), buffer_next(ob, valid.numpy()))
    np.random.seed(5)
    samp = buf.sample_indices(64).numpy()
    assert np.isin(samp, ob.sample_all()).all()
    buf.reset()
    assert len(buf) == 0 and buf.sample_indices(0).numel() == 0


def test_batch_container():
    from fsrl_b200.data import Batch
    b = Batch(obs=np.arange(6).reshape(3, 2), info={"cost": np.array([0., 1., 0.])}, policy=Batch())
    assert len(b) == 3 and b.info.cost[1] == 1.0
    s = b[1:]
    assert len(s) == 2 and s.info.cost.tolist() == [1.0, 0.0]
    b.update(act=np.zeros(3))
    assert "act" in b and b.get("missing", 7) == 7 and b.pop("act").shape == (3,)


def test_config_tables_match_reference_defaults():
    from fsrl_b200.config import cpo_cfg, ddpgl_cfg, ppol_cfg, sacl_cfg
    p = ppol_cfg.TrainCfg()
    assert (p.lr, p.target_kl, p.vf_coef, p.max_grad_norm, p.eps_clip, p.batch_size, p.repeat_per_collect,
            p.episode_per_collect, p.hidden_sizes, p.lagrangian_pid) == \
        (5e-4, 0.02, 0.25, 0.5, 0.2, 256, 4, 20, (128, 128), (0.05, 0.0005, 0.1))       # ppol_cfg.py:14-49
    c = cpo_cfg.TrainCfg()
    assert (c.lr, c.target_kl, c.max_backtracks, c.optim_critic_iters, c.l2_reg, c.batch_size) == \
        (1e-3, 0.01, 100, 10, 0.001, 99999)                                                 # cpo_cfg.py:14-45
    s = sacl_cfg.TrainCfg()
    assert (s.actor_lr, s.critic_lr, s.alpha_lr, s.tau, s.n_step, s.gamma, s.update_per_step) == \
        (5e-4, 1e-3, 3e-4, 0.05, 2, 0.97, 0.2)
    d = ddpgl_cfg.MujocoBaseCfg()
    assert (d.gamma, d.n_step, d.buffer_size, d.cost_limit, d.exploration_noise) == (0.99, 3, 800000, 25, 0.1)
    assert ppol_cfg.Bullet10MCfg().epoch == 1000 and dataclasses.is_dataclass(p)


def test_compat_shims_resolve_reference_imports():
    import fsrl_b200.compat as compat
    compat.install()
    import gymnasium as gym
    from tianshou.env import ShmemVectorEnv
    from tianshou.utils.net.continuous import ActorProb
    from fsrl.agent import PPOLagAgent
    from fsrl.policy import PPOLagrangian
    from fsrl.utils.exp_util import auto_name
    env = gym.make("SafetyCarCircle-v0")
    assert env.observation_space.shape == (8,) and env.action_space.shape == (2,)
    assert env.spec.max_episode_steps == 300
    name = auto_name({"a": 1, "b": 2}, {"a": 1, "b": 3}, "ppol")       # reference format: prefix_<key><value>-<tag>
    assert name[:-5] == "ppol_b3" and name[-5] == "-" and len(name) == len("ppol_b3") + 5
    with pytest.raises(KeyError):
        gym.make("NoSuchTask-v0")


def test_logger_formats(tmp_path):
    from fsrl_b200.utils.exp_util import load_config_and_model
    from fsrl_b200.utils.logger import BaseLogger
    lg = BaseLogger(str(tmp_path), log_txt=True, name="run")
    lg.save_config({"task": "SafetyCarCircle-v0", "hidden_sizes": (64, 64)}, verbose=False)
    lg.setup_checkpoint_fn(lambda: {"model": {"w": torch.ones(2)}})
    lg.store(tab="train", reward=1.0); lg.store(tab="train", reward=3.0)
    lg.store_many("loss", "kl", [0.1, 0.3])
    assert lg.get_mean("train/reward") == 2.0 and abs(lg.get_mean("loss/kl") - 0.2) < 1e-12
    lg.save_checkpoint()
    lg.write(100)
    lines = open(os.path.join(tmp_path, "run", "progress.txt")).read().strip().split("\n")
    assert lines[0].split("\t")[0] == "Steps" and lines[1].split("\t")[0] == "100"
    cfg, model = load_config_and_model(os.path.join(tmp_path, "run"))
    assert cfg["task"] == "SafetyCarCircle-v0" and torch.equal(model["model"]["w"], torch.ones(2))


def test_tensorboard_logger_round_trip(tmp_path):
    """TensorboardLogger writes scalars per key and restores (epoch, env_step, gradient_step) from
    its own event file (tb_logger.py:45-81: resume=True reads them back)."""
    pytest.importorskip("tensorboard")
    from fsrl_b200.utils.logger import TensorboardLogger
    lg = TensorboardLogger(str(tmp_path), log_txt=True, name="tb_run")
    for epoch, (env_step, grad_step) in enumerate([(1000, 40), (2000, 80)], start=1):
        lg.store(tab="update", episode=epoch, gradient_steps=grad_step)
        lg.store(tab="train", reward=float(epoch))
        lg.write(env_step, display=False)
    lg.summary_writer.close()
    lg2 = TensorboardLogger(str(tmp_path), log_txt=False, name="tb_run")
    epoch, env_step, gradient_step = lg2.restore_data()
    # the reference reads the STEP of the last item (the x axis = env steps), tb_logger.py:66-77
    assert (epoch, env_step, gradient_step) == (2000, 2000, 2000)
    lines = open(os.path.join(tmp_path, "tb_run", "progress.txt")).read().strip().split("\n")
    assert lines[0].split("\t")[0] == "Steps" and len(lines) == 3


def test_env_registry_and_dims_without_gpu():
    from fsrl_b200 import envs
    for task, (D, A, T) in {"SafetyCarCircle-v0": (8, 2, 300), "SafetyCarRun-v0": (7, 2, 200),
                            "SafetyBallCircle-v0": (8, 2, 200), "SafetyBallRun-v0": (7, 2, 100),
                            "SafetyAntCircle-v0": (34, 8, 500), "SafetyPointGoal1Gymnasium-v0": (60, 2, 1000)}.items():
        e = envs.make(task)
        assert e.observation_space.shape == (D,) and e.action_space.shape == (A,) and e.spec.max_episode_steps == T


def test_running_mean_std_matches_batch_statistics():
    """Chunked updates reproduce the moments of the concatenated stream (parallel-variance update)."""
    from fsrl_b200.utils.optim_util import RunningMeanStd
    rng = np.random.default_rng(0)
    xs = [rng.normal(3.0, 2.0, size=n) for n in (1, 7, 300, 4096)]
    rms = RunningMeanStd()
    for x in xs:
        rms.update(x)
    full = np.concatenate(xs)
    assert rms.count == len(full)
    assert rms.mean == pytest.approx(full.mean(), rel=1e-12) and rms.var == pytest.approx(full.var(), rel=1e-10)


def test_oracle_offpolicy_step_equals_targets_plus_update():
    """oracle glue check (CPU): sac_step / ddpg_step == n-step targets from the buffer followed by
    sac_update / ddpg_update (the halves pinned separately: nstep_return by the reference's numba
    function, the updates by the reference's own learn())."""
    import copy
    from oracle import nets as onets, offpolicy as ooff
    from oracle.collector import OracleBuffer
    rng = np.random.default_rng(2)
    D, A, H, E, T = 5, 2, 8, 3, 12
    buf = OracleBuffer(E * T, E, D, A)
    for t in range(T):
        ids = np.arange(E)
        buf.add(ids, rng.normal(size=(E, D)).astype(np.float32), np.tanh(rng.normal(size=(E, A))).astype(np.float32),
                rng.normal(size=E).astype(np.float32), (rng.random(E) < 0.2).astype(np.float32), np.zeros(E, np.float32),
                rng.random(E) < 0.1, np.full(E, t == T - 1), rng.normal(size=(E, D)).astype(np.float32))
    idx = rng.integers(0, E * T, size=16).astype(np.int64)
    torch.manual_seed(0)
    actor = onets.GaussActor(D, A, [H, H], unbounded=True, conditioned_sigma=True)
    crit = [[onets.ValueNet(D + A, [H, H]) for _ in range(2)] for _ in range(2)]
    crit_old = copy.deepcopy(crit)
    eps_n, eps_c = torch.randn(16, A), torch.randn(16, A)

    def run(step_fn):
        a, c, co = copy.deepcopy(actor), copy.deepcopy(crit), copy.deepcopy(crit_old)
        ao = torch.optim.Adam(a.parameters(), lr=1e-3)
        cop = torch.optim.Adam([p for pair in c for q in pair for p in q.parameters()], lr=1e-3)
        return step_fn(a, c, co, ao, cop), a

    (st1, _), a1 = run(lambda a, c, co, ao, cop: ooff.sac_step(a, c, co, ao, cop, buf, idx, eps_n, eps_c, alpha=0.2, gamma=0.97,
                                                                n_step=2, tau=0.05, lagrangian=0.4))

    def manual(a, c, co, ao, cop):
        with torch.no_grad():
            _, terminal = ooff.nstep_targets(buf, idx, [np.zeros(16)] * 2, 0.97, 2)
            on = torch.from_numpy(buf.obs_next[terminal])
            an, lpn = ooff.sac_forward(a, on, eps_n)
            tq = [(torch.min(co[i][0](on, an), co[i][1](on, an)) - 0.2 * lpn).numpy() for i in range(2)]
        rets, _ = ooff.nstep_targets(buf, idx, tq, 0.97, 2)
        return ooff.sac_update(a, c, co, ao, cop, torch.from_numpy(buf.obs[idx]), torch.from_numpy(buf.act[idx]),
                               torch.from_numpy(rets), eps_c, alpha=0.2, tau=0.05, lagrangian=0.4)

    (st2, _), a2 = run(manual)
    assert st1 == st2
    for p, q in zip(a1.parameters(), a2.parameters()):
        assert torch.equal(p, q)
    # DDPG
    dact = onets.DetActor(D, A, [H, H])
    dcrit = [onets.ValueNet(D + A, [H, H]) for _ in range(2)]
    a, ao_, c, co = copy.deepcopy(dact), copy.deepcopy(dact), copy.deepcopy(dcrit), copy.deepcopy(dcrit)
    st = ooff.ddpg_step(a, ao_, c, co, torch.optim.Adam(a.parameters(), lr=1e-3),
                        torch.optim.Adam([p for q in c for p in q.parameters()], lr=1e-3), buf, idx, gamma=0.97, n_step=2,
                        tau=0.05, lagrangian=0.4)
    assert np.isfinite(list(st.values())).all() and {"loss/q0", "loss/q1", "loss/actor_total"} <= set(st)


def test_speculative_permutation_keeps_the_numpy_stream():
    """ppo_lag.learn draws the next learn call's first permutation early (while the last launch runs): the global NumPy
    stream must look untouched to everybody else, and the draw must be dropped if someone consumed the stream in between."""
    from fsrl_b200.policy.ppo_lag import PPOLagrangian as P

    class Host:
        _same_rng_state = staticmethod(P._same_rng_state)

    h = Host()
    np.random.seed(11)
    ref = [np.random.permutation(50) for _ in range(3)]
    np.random.seed(11)
    first = np.random.permutation(50)
    P._prefetch_permutation(h, 50)
    second = P._first_permutation(h, 50)
    third = np.random.permutation(50)
    assert (first == ref[0]).all() and (second == ref[1]).all() and (third == ref[2]).all()
    # somebody else draws between the two learn calls: the reference order is (their draw, then the permutation)
    np.random.seed(11)
    r_ref = np.random.rand(); p_ref = np.random.permutation(50)
    np.random.seed(11)
    P._prefetch_permutation(h, 50)
    r = np.random.rand()
    p = P._first_permutation(h, 50)
    assert r == r_ref and (p == p_ref).all()
    # a different row count invalidates the draw
    np.random.seed(11)
    P._prefetch_permutation(h, 50)
    q = P._first_permutation(h, 40)
    np.random.seed(11)
    assert (q == np.random.permutation(40)).all()
