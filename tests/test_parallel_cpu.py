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


def _tr_worker(rank, world, port, q):
    """TrustRegionMixin's data-parallel combiners on CPU tensors over gloo: weighted vectors,
    global batch sums and global advantage standardisation must equal the single-process result on
    the union of the ranks' (unequally sized) batches."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fsrl_b200.parallel import DataParallel
    from fsrl_b200.policy.trust_region import TrustRegionMixin

    class Host(TrustRegionMixin):
        device = torch.device("cpu")

    h = Host()
    h._dp = DataParallel(dist, with_nccl=False)
    rng = np.random.default_rng(5)
    sizes = [700, 1300]
    data = [rng.normal(2.0, 3.0, size=n).astype(np.float32) for n in sizes]      # same on both ranks
    grads = [rng.normal(size=16).astype(np.float32) for _ in sizes]              # per-rank local MEAN gradients
    n = sizes[rank]
    n_g = h._dp_begin(n)
    v = torch.from_numpy(grads[rank].copy())
    h._gvec(v)                                                                    # -> global-batch mean gradient
    h._sums = torch.tensor([float(data[rank].sum()), float(n), 0.0, 0.0], dtype=torch.float64)
    sums = h._gsums()
    x = torch.from_numpy(data[rank].copy())
    h._standardize(x, n)
    h._dp_same_count(3, "minibatch count")                                       # equal counts: no error
    try:
        h._dp_same_count(3 + rank, "minibatch count")
        mismatch_raises = False
    except RuntimeError:
        mismatch_raises = True
    q.put((rank, n_g, v.numpy(), sums, x.numpy(), mismatch_raises))
    dist.destroy_process_group()


def test_trust_region_combiners_equal_union_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() + 777) % 2000)
    procs = [ctx.Process(target=_tr_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    sizes = [700, 1300]
    data = [rng.normal(2.0, 3.0, size=n).astype(np.float32) for n in sizes]
    grads = [rng.normal(size=16).astype(np.float32) for _ in sizes]
    union = np.concatenate(data).astype(np.float64)
    want_g = (sizes[0] * grads[0].astype(np.float64) + sizes[1] * grads[1]) / sum(sizes)
    for rank, n_g, v, sums, x, mismatch_raises in res:
        assert n_g == 2000
        np.testing.assert_allclose(v, want_g, rtol=1e-6)
        assert sums[0] == pytest.approx(union.sum(), rel=1e-6) and sums[1] == 2000
        want_x = (data[rank] - union.mean()) / union.std(ddof=1)
        np.testing.assert_allclose(x, want_x, rtol=2e-5, atol=2e-6)
        assert mismatch_raises
    np.testing.assert_array_equal(res[0][2], res[1][2])      # identical reduced vector on both ranks


def _trainer_worker(rank, world, port, q):
    """BaseTrainer's cycle loop with ranks that collect DIFFERENT step counts per cycle (early terminations on one rank):
    every rank must leave the loop after the same number of cycles and size its updates from the same agreed count."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fsrl_b200.parallel import DataParallel
    from fsrl_b200.trainer.base_trainer import BaseTrainer
    from fsrl_b200.utils.logger import DummyLogger

    dp = DataParallel(dist, with_nccl=False)

    class Policy:
        _dp = dp
        def train(self): pass
        def eval(self): pass

    class Collector:
        collect_time, collect_step = 1e-3, 0
        def reset_stat(self): pass
        def collect(self, n_episode):
            n_st = 300 if rank == 0 else 240          # rank 1's episodes end early
            self.collect_step += n_st
            local = {"n/ep": 1, "n/st": n_st, "rew": 1.0, "len": float(n_st), "total_cost": 2.0, "cost": 2.0,
                     "truncated": 1.0, "terminated": 0.0}
            return dp.reduce_collect_stats(local)     # what parallel.attach's pre_update_fn hook feeds the policy

    class Trainer(BaseTrainer):
        cycles, sizes = 0, []
        def policy_update_fn(self, stats_train):
            self.cycles += 1
            self.sizes.append(self._cycle_steps)

    tr = Trainer("onpolicy", Policy(), Collector(), None, max_epoch=2, step_per_epoch=1000, episode_per_test=1,
                 episode_per_collect=1, logger=DummyLogger(), verbose=False, show_progress=False)
    tr.run()
    q.put((rank, tr.cycles, tuple(tr.sizes), tr.env_step))
    dist.destroy_process_group()


def test_trainer_cycles_agree_when_ranks_collect_different_step_counts():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, c0, s0, e0), (_, c1, s1, e1) = res
    assert c0 == c1 == 8                                  # 2 epochs x ceil(1000 / 300): the quota advances by the agreed count
    assert s0 == s1 == (300,) * 8                         # the larger rank's count, on both ranks
    assert e0 == e1 == 8 * 540                            # env_step counts the GLOBAL steps the reduced statistics report
