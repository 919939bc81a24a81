"""Data-parallel plumbing: one process per GPU, every rank owns E/G envs, its
own rollout buffer and computes GAE locally (segments never cross envs).  Shared state is kept
identical on all ranks by

* per collect : one all-reduce of [sum cost, episodes, sum return, sum length] -> identical PID
                multiplier everywhere (the reference's ``stats_train["cost"]`` becomes the global
                mean episodic cost);
* per repeat  : one all-reduce of the per-minibatch advantage moments, so the per-minibatch
                normalisation (ppo_lag.py:178-182) is over the GLOBAL minibatch;
* per step    : the flat gradient buffers are exchanged over PEER MEMORY (CUDA-IPC mapped blocks,
                NVLink): ppo_dp_reduce_kernel signals, waits and sums all ranks' buffers in rank
                order, fused with the gradient norm -- no NCCL launch on the step path (NCCL
                all-reduce remains as the fallback when peer mapping is unavailable);
* per repeat  : the KL early-stop statistic is averaged so that all ranks stop together.

Off-policy learners (SAC / DDPG): every rank samples its own replay shard; the C loop
(csrc/offpolicy.cu) all-reduces the critics' and the actor's gradient slices once per gradient
step and the entropy-tuning statistic, so parameters, targets and alpha stay identical.
Trust-region learners (CPO / TRPO): gradient vectors, every Hessian-vector product and the
batch sums of the line search are combined with weights n_r / sum(n) (policy/trust_region.py).

The reference has no distributed code; this file is the new engine's design.
Host-side scalar reductions go through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests); gradients go through our own NCCL communicator created from a broadcast unique id.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch


class DataParallel:
    def __init__(self, dist, device: Optional[torch.device] = None, with_nccl: bool = True):
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.comm = None
        if with_nccl and self.world > 1:
            from . import _lib
            buf = torch.zeros(128, dtype=torch.uint8, device=self.device)
            if self.rank == 0:
                raw = ctypes.create_string_buffer(128)
                _lib.check(_lib.lib.fsrl_comm_unique_id(raw))
                buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
            dist.broadcast(buf, src=0)
            idb = bytes(buf.cpu().numpy().tobytes())
            comm = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib.fsrl_comm_init(idb, self.rank, self.world, ctypes.byref(comm)))
            self.comm = comm

    # ---- peer-memory gradient exchange (csrc/ppo.cu::ppo_dp_reduce_kernel) -----------------------
    def enable_p2p(self, n_floats: int) -> bool:
        """Allocate this rank's exchange block, swap CUDA-IPC handles with the peers and map their
        blocks.  Returns False (NCCL stays in use) if the world is too large or mapping fails."""
        from . import _lib
        lib = _lib.lib
        if self.world < 2 or self.world > 8 or n_floats > 4096 * 1024 or getattr(self, "p2p", None) is not None:
            return getattr(self, "p2p", None) is not None
        with torch.cuda.device(self.device):
            base = ctypes.c_void_p()
            raw = ctypes.create_string_buffer(64)
            _lib.check(lib.fsrl_p2p_alloc(int(n_floats), ctypes.byref(base), raw))
            mine = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).to(self.device)
            allh = [torch.zeros(64, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
            self.dist.all_gather(allh, mine)
            bases, ok = [], 1
            for r in range(self.world):
                if r == self.rank:
                    bases.append(base.value)
                    continue
                pb = ctypes.c_void_p()
                rc = lib.fsrl_p2p_open(bytes(allh[r].cpu().numpy().tobytes()), ctypes.byref(pb))
                if rc != 0:
                    ok = 0
                bases.append(pb.value)
            ok = int(self.all_max([-ok])[0]) == -1          # everybody mapped everybody?
        if not ok:
            return False
        stride = int(lib.fsrl_p2p_stride(int(n_floats)))
        self.p2p = {"bases": bases, "n": int(n_floats), "stride": stride,
                    "xg": [[b + par * stride * 4 for b in bases] for par in (0, 1)],
                    "flags": [b + 2 * stride * 4 for b in bases],
                    "err": base.value + 2 * stride * 4 + 8 * 8,
                    "part": base.value + 2 * stride * 4 + 8 * 8 + 64}
        return True

    def fill_p2p(self, u) -> None:
        """Hand the mapped exchange blocks to an update descriptor (fsrl_ppo_update_t)."""
        p = getattr(self, "p2p", None)
        if p is None:
            u.p2p_on = 0
            return
        for par in (0, 1):
            for r in range(self.world):
                u.p2p_xg[par][r] = p["xg"][par][r]
        for r in range(self.world):
            u.p2p_flags[r] = p["flags"][r]
        u.p2p_err, u.p2p_part, u.p2p_rank, u.p2p_on = p["err"], p["part"], self.rank, 1
        u.p2p_stride = p["stride"]

    def p2p_check(self) -> None:
        p = getattr(self, "p2p", None)
        if p is None:
            return
        from . import _lib
        err = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.fsrl_p2p_poll_error(p["err"], ctypes.byref(err)))
        if err.value:
            raise RuntimeError("data-parallel gradient exchange: a peer rank never arrived (timed out)")

    # ---- host-side scalar reductions -----------------------------------------------------------
    def all_sum(self, values) -> np.ndarray:
        t = torch.as_tensor(np.asarray(values, dtype=np.float64), device=self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def all_max(self, values) -> np.ndarray:
        t = torch.as_tensor(np.asarray(values, dtype=np.float64), device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def reduce_collect_stats(self, stats: Dict) -> Dict:
        """Global view of one collect: episodic means over ALL ranks' episodes."""
        n_ep, n_st = stats["n/ep"], stats["n/st"]
        loc = [stats["total_cost"], n_ep, stats["rew"] * n_ep, stats["len"] * n_ep, n_st,
               stats["truncated"] * n_ep, stats["terminated"] * n_ep]
        g = self.all_sum(loc)
        ep = max(g[1], 1.0)
        out = dict(stats)
        out.update({"total_cost": g[0], "n/ep": int(g[1]), "n/st": int(g[4]), "cost": g[0] / ep,
                    "rew": g[2] / ep, "len": g[3] / ep, "truncated": g[5] / ep, "terminated": g[6] / ep,
                    "local": stats})
        return out

    def mean_scalar(self, x: float) -> float:
        return float(self.all_sum([x])[0] / self.world)

    def broadcast_(self, t: torch.Tensor, src: int = 0) -> None:
        self.dist.broadcast(t, src=src)


def shard_seed(seed: int, rank: int) -> int:
    """Independent env / noise streams per rank (same network init seed everywhere)."""
    return (int(seed) + 1000003 * int(rank)) & 0xFFFFFFFF


def attach(policy, dist, device=None, p2p: bool = True) -> DataParallel:
    """Make `policy` data parallel: broadcast rank 0's parameters, hook the collect-statistics
    reduction into pre_update_fn, and hand the NCCL communicator to the update descriptor."""
    device = device if device is not None else policy.device
    if type(policy).__name__ == "FOCOPS":
        raise NotImplementedError("FOCOPS has no data-parallel update yet (SURVEY.md 8e covers PPO-Lag, CPO, "
                                  "TRPO-Lag, SAC-Lag, DDPG-Lag)")
    dp = DataParallel(dist, device)
    dp.broadcast_(policy.arena.theta)
    if hasattr(policy, "_mirror_dirty"):
        policy._mirror_dirty = True
    policy._dp = dp
    from .policy.ppo_lag import PPOLagrangian
    if p2p and dp.world > 1 and isinstance(policy, PPOLagrangian):
        # PPO-Lag: gradients over peer memory; the buffers also hold the per-CTA gradient tiles / flags of the
        # persistent update kernel (csrc/ppo_persist.cu)
        from . import _lib
        need = int(_lib.lib.fsrl_ppo_persist_p2p_floats(len(policy.arena.slots)))
        dp.enable_p2p(max(policy.arena.theta.numel(), need))
    if hasattr(policy, "_upd_seed"):      # off-policy rsample / exploration noise: one stream per rank
        policy._upd_seed = shard_seed(policy._upd_seed, dp.rank)
    inner = policy.pre_update_fn

    def pre_update_fn(stats_train, **kw):
        return inner(stats_train=dp.reduce_collect_stats(stats_train), **kw)

    policy.pre_update_fn = pre_update_fn
    return dp
