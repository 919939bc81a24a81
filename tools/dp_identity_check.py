"""Lock-step check of the persistent PPO kernel's data-parallel exchange: run launches of growing length on every rank
and compare the ranks' parameters bit by bit after each one.
Run:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_identity_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def groups(policy):
    out = []
    for n, s in enumerate(policy.arena.slots):
        o = s.offset
        for name, cnt in (("W1", s.D * s.H), ("b1", s.H), ("W2", s.H * s.H), ("b2", s.H), ("W3", s.H * s.out), ("b3", s.out)):
            out.append(("net%d.%s" % (n, name), o, o + cnt))
            o += cnt
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(device))
    from helpers import build_ppo
    from test_ppo_scale_gpu import _sub_batch
    from fsrl_b200 import parallel
    policy, venv, buf, col = build_ppo("SafetyCarCircle-v0", hidden=(256, 256), n_env=1024, seed=10, device=device, max_grad_norm=0.5)
    venv.seed(parallel.shard_seed(12, rank)); col.reset_env()
    policy.set_action_seed(parallel.shard_seed(11, rank))
    dp = parallel.attach(policy, dist, device=device, p2p=True)
    stats = col.collect(n_episode=1024)
    policy.pre_update_fn(stats_train=stats)
    full = policy.process_fn(None, buf, buf.sample_indices(0))
    policy._target_kl = 1e9
    gr = groups(policy)

    def compare(label):
        th = policy.arena.theta.detach().contiguous()
        allth = [torch.empty_like(th) for _ in range(world)]
        dist.all_gather(allth, th)
        if rank == 0:
            bad_total = 0
            for r in range(1, world):
                neq = allth[0].view(torch.int32) != allth[r].view(torch.int32)
                bad_total += int(neq.sum())
                if neq.any():
                    parts = ["%s:%d" % (g, int(neq[a:b].sum())) for g, a, b in gr if neq[a:b].any()]
                    idx = neq.nonzero().flatten()[:12].tolist()
                    print("  %s: rank %d differs in %d parameters (max |d| %.2e): %s; first %s" % (
                        label, r, int(neq.sum()), float((allth[0] - allth[r]).abs().max()), " ".join(parts), idx), flush=True)
            if bad_total == 0:
                print("  %s: identical on all %d ranks" % (label, world), flush=True)
        # re-align parameters AND Adam moments so that later launches start identical again
        for t in (policy.arena.theta, policy.optim.m, policy.optim.v):
            dist.broadcast(t, src=0)
        policy._mirror_dirty = True

    for n_steps, reps in ((8, 1), (512, 2), (2400, 12)):
        batch = _sub_batch(policy, full, n_steps * 256)
        for i in range(reps):
            np.random.seed(300 + i)
            policy.learn(batch, batch_size=256, repeat=1)
            torch.cuda.synchronize()
            compare("launch of %4d steps, #%d" % (n_steps, i + 1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
