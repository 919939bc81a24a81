"""world_size-2 gloo tests of the data-parallel host logic: collect-statistics
reduction -> identical PID multiplier on every rank, KL agreement, seed sharding.  The gradient
all-reduce itself runs over NCCL on the GPUs (tests/test_parallel_gpu.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fsrl_b200.parallel import DataParallel, shard_seed
    from fsrl_b200.utils.optim_util import LagrangianOptimizer
    dp = DataParallel(dist, with_nccl=False)
    # each rank saw different episodes
    stats = {"n/ep": 3 + rank, "n/st": 900 + 300 * rank, "rew": 10.0 * (rank + 1), "len": 300.0,
             "total_cost": 30.0 + 12 * rank, "cost": (30.0 + 12 * rank) / (3 + rank), "truncated": 1.0,
             "terminated": 0.0}
    g = dp.reduce_collect_stats(stats)
    pid = LagrangianOptimizer((0.05, 0.0005, 0.1))
    pid.step(g["cost"], 10.0)
    kl = dp.mean_scalar(0.01 * (rank + 1))
    q.put((rank, g["cost"], g["n/ep"], g["n/st"], g["rew"], pid.get_lag(), kl, shard_seed(10, rank)))
    dist.destroy_process_group()


def test_collect_stats_and_pid_agree_across_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, c0, ep0, st0, r0, l0, k0, s0), (_, c1, ep1, st1, r1, l1, k1, s1) = res
    assert c0 == c1 == pytest.approx((30.0 + 42.0) / 7) and ep0 == ep1 == 7 and st0 == st1 == 2100
    assert r0 == r1 == pytest.approx((10.0 * 3 + 20.0 * 4) / 7)
    assert l0 == l1 and l0 > 0                       # identical dual variable on both ranks
    assert k0 == k1 == pytest.approx(0.015)
    assert s0 != s1                                  # but independent env / noise streams
