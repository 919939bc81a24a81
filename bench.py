"""bench.py -- BASELINE.json's metric on its headline config.

metric : env-steps/sec over collect+update (the reference's ``train_speed``,
         fsrl/trainer/base_trainer.py:345-347)
config : c2 = PPO-Lagrangian, SafetyCarCircle-v0, 2048 envs, 2x256 MLP, batch_size 256,
         repeat_per_collect 4, episode_per_collect = 2048 (one 300-step episode per env and
         collect => 614 400 transitions per step), fp32
step   : ONE collect + update cycle (OnpolicyTrainer.train_step + policy_update_fn)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints one JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is derived.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASK = "SafetyCarCircle-v0"
ENVS = 2048
HIDDEN = (256, 256)
BATCH = 256
REPEAT = 4
SEED = 10


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def build(device, rank, envs=ENVS, hidden=HIDDEN):
    from fsrl_b200 import envs as fenvs
    from fsrl_b200.agent import PPOLagAgent
    from fsrl_b200.data import FastCollector, VectorReplayBuffer
    from fsrl_b200.trainer import OnpolicyTrainer
    from fsrl_b200.utils.logger import BaseLogger
    demo = fenvs.make(TASK)
    logger = BaseLogger()
    agent = PPOLagAgent(demo, logger=logger, cost_limit=10, device=device, seed=SEED, lr=5e-4,
                        hidden_sizes=hidden, max_grad_norm=0.5)      # cfg values (ppol_cfg.py:14-33)
    T = demo.spec.max_episode_steps
    from fsrl_b200.parallel import shard_seed
    train_envs = fenvs.DeviceVectorEnv(TASK, envs, device=device, seed=shard_seed(SEED + 1, rank))
    agent.policy.set_action_seed(shard_seed(SEED + 7, rank))
    # fixed work per step: the KL early stop (ppo_lag.py:251-255) is disabled so that EVERY step
    # runs all `REPEAT` passes = 4 x 2400 minibatch updates (the most work the config can do)
    agent.policy._target_kl = float("inf")
    buf = VectorReplayBuffer(envs * T, envs, device=device)
    col = FastCollector(agent.policy, train_envs, buf, exploration_noise=True)
    trainer = OnpolicyTrainer(agent.policy, col, None, max_epoch=1, batch_size=BATCH, cost_limit=10,
                              step_per_epoch=envs * T, repeat_per_collect=REPEAT,
                              episode_per_collect=envs, episode_per_test=1, logger=logger,
                              verbose=False, show_progress=False)
    return agent, trainer, col, buf, T


def one_cycle(trainer):
    stats = trainer.train_step()
    trainer.policy_update_fn(stats)
    return stats


def phase_times(agent, col, buf, iters=200):
    """Average duration of the dominant update kernels, CUDA events on the launching stream
    (fsrl_ppo_phase_times in csrc/ppo.cu launches each phase kernel `iters` times back to back
    on a real 256-row minibatch of the batch that was just trained on)."""
    import ctypes
    from fsrl_b200 import _lib
    pol = agent.policy
    col.collect(ENVS)                      # the trainer resets the buffer after every update
    idx = buf.sample_indices(0)
    batch = pol.process_fn(None, buf, idx)
    n = batch.n
    perm = torch.randperm(n, device=pol.device).to(torch.int32)
    pol._ensure_update_state(BATCH, n, 1)
    u = pol._descriptor(batch, perm)
    ms = (ctypes.c_float * 4)()
    _lib.check(_lib.lib.fsrl_ppo_phase_times(ctypes.byref(u), BATCH, iters, ms, torch.cuda.current_stream().cuda_stream))
    return [float(x) for x in ms]


def gae_time(buf, policy, iters=20):
    from fsrl_b200 import ops
    n = buf.maxsize
    v = torch.randn(2, n, device=policy.device); vn = torch.randn(2, n, device=policy.device)
    end = (buf.terminated | buf.truncated)
    adv = torch.empty_like(v); ret = torch.empty_like(v)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=policy.device)
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); ops.gae_dual(v, vn, buf.rew, buf.cost, end, buf.terminated, 0.99, 0.95, out=(adv, ret)); e.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(e))
    return float(np.mean(ts)), n


def gae_time_large(device, envs=ENVS * 16, T=300, iters=10):
    """The same kernel on a 16x larger synthetic collect (SURVEY.md 8d asks for large-E sweeps where
    the HBM bound is reachable): 412 MB of algorithmic traffic, larger than L2, flushed anyway."""
    from fsrl_b200 import ops
    from fsrl_b200.utils.synth import synth_gae_inputs
    d = synth_gae_inputs(envs, T, seed=10)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    v, vn, r, c = dev(d["v"]), dev(d["vnext"]), dev(d["rew"]), dev(d["cost"])
    end = dev((d["terminated"] | d["truncated"]).astype(np.uint8))
    term = dev(d["terminated"].astype(np.uint8))
    adv = torch.empty_like(v); ret = torch.empty_like(v)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); ops.gae_dual(v, vn, r, c, end, term, 0.99, 0.95, out=(adv, ret)); e.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(e))
    return float(np.mean(ts)), envs * T


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))
    agent, trainer, col, buf, T = build(device, rank)
    if world > 1:
        from fsrl_b200 import parallel
        parallel.attach(agent.policy, dist, device=device)
    steps_per_cycle = ENVS * T

    for _ in range(args.warmup):
        one_cycle(trainer)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-timed region: K collect+update cycles ---------------------------------------------
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    from fsrl_b200 import _lib as _fl
    launches0 = int(_fl.lib.fsrl_launch_count())
    ev0.record()
    collect_s = 0.0
    for _ in range(args.steps):
        c0 = col.collect_time
        one_cycle(trainer)
        collect_s += col.collect_time - c0
    ev1.record()
    torch.cuda.synchronize()
    t_wall = time.time() - t_wall0
    launches = int(_fl.lib.fsrl_launch_count()) - launches0
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    total_steps = steps_per_cycle * args.steps * world
    value = total_steps / (ms * 1e-3)
    # ---- e2e: a SECOND region of K cycles through the public trainer API, wall clock, synchronised on
    # both sides.  Every cycle uploads the minibatch permutations from pinned host memory and downloads
    # the per-minibatch statistics + collect statistics the trainer logs; observations never exist on
    # the host (the environment model runs on the device), so these ARE the path's copies.
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_e2e0 = time.time()
    for _ in range(args.steps):
        one_cycle(trainer)
    torch.cuda.synchronize()
    t_e2e = time.time() - t_e2e0
    if dist is not None:
        t = torch.tensor([t_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- e2e: the same cycles through the public trainer API, wall-clock incl. every host<->device
    # copy the API performs (minibatch permutations up, statistics down) ----------------------------
    n_mb = (steps_per_cycle + BATCH - 1) // BATCH
    h2d = REPEAT * steps_per_cycle * 4                           # int32 permutation per repeat
    d2h = REPEAT * n_mb * 8 * 4 + 64                             # per-minibatch stats + collect stats
    e2e_value = total_steps / t_e2e
    # ---- roofline of the dominant kernel + GAE ----------------------------------------------------
    pk, how = peaks()
    ph = phase_times(agent, col, buf)
    D, A, H = 8, 2, HIDDEN[0]
    fl_net = lambda out: 2 * BATCH * (D * H + H * H + H * out) + 2 * BATCH * (H * out + H * H)
    flops_a = fl_net(A) + 2 * fl_net(1)
    ach = flops_a / ((ph[0] + ph[1]) * 1e-3) / 1e12
    gms, gn = gae_time(buf, agent.policy)
    gae_bytes = gn * 42
    gms_l, gn_l = gae_time_large(device)
    out = {
        "metric": "env-steps/sec (collect+update) SafetyCarCircle-v0 PPO-Lag",
        "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: PPO-Lagrangian SafetyCarCircle-v0, 2048 envs/GPU x 300 steps, "
                               "2x256 MLP, batch_size 256, repeat 4, max_grad_norm 0.5 "
                               "(analytic on-device env model, random-init weights)",
                   "envs_per_gpu": ENVS, "transitions_per_step": steps_per_cycle * world,
                   "parallelism": f"dp{world}",
                   "kl_early_stop": "disabled (fixed work: 4 repeats x 2400 minibatch updates per step)",
                   "l2": "working set per cycle (buffers 53 MB + per-minibatch gathers over 614k rows) "
                         "cycles through > L2-size of distinct data between reuses; no explicit flush"},
        "collect_s_per_step": collect_s / args.steps,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "how": "separate region: K more collect+update cycles through OnpolicyTrainer.train_step / "
                       "policy_update_fn, wall clock, max over ranks; H2D = pinned minibatch permutations, "
                       "D2H = per-minibatch + collect statistics"},
        "gpu_launches": launches,
        "roofline": {"kernel": "ppo_fwd_kernel<256> + ppo_bwd_kernel<256>", "bound": "tensor", "achieved": ach,
                     "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                     "traffic": 3.44e6, "traffic_unit": "DRAM bytes per fwd+bwd launch pair (ncu --set full, "
                                                        "profiles/r1_ppo_update_ncu_summary.txt): the working set is L2 resident",
                     "peak_source": how + " bf16 burst",
                     "note": "fp32-faithful 3xTF32 split-operand mma.sync (legacy tensor path, 3 MMAs per fp32 product); flops = algorithmic fwd+bwd of 3 MLPs on a 256-row minibatch; latency-bound (9600 dependent optimiser steps of 256 rows)",
                     "phase_ms": {"fwd": ph[0], "bwd": ph[1], "wgrad": ph[2], "adam": ph[3]}},
        "roofline_gae": {"kernel": "gae_dual_kernel<2,true>", "bound": "hbm",
                         "achieved": gae_bytes / (gms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": gae_bytes / (gms * 1e-3) / 1e9 / pk["hbm_gbs"], "ms": gms,
                         "bytes_per_transition": 42, "peak_source": how, "transitions": gn,
                         "traffic": 16.0e6 + 9.8e6, "traffic_unit": "DRAM bytes per launch on the c2 collect: 16.0 MB read "
                         "(ncu, = algorithmic 26 B/transition) + 9.8 MB of adv/ret written back from L2 after the kernel",
                         "note": "c2-sized collect (25.8 MB): launch + one latency chain per tile dominate",
                         "large": {"transitions": gn_l, "ms": gms_l,
                                   "achieved": gn_l * 42 / (gms_l * 1e-3) / 1e9,
                                   "frac": gn_l * 42 / (gms_l * 1e-3) / 1e9 / pk["hbm_gbs"]}},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_reference(sample_envs=args.cpu_envs, cycles=1)
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle restatement of the reference path on the host cores
# -------------------------------------------------------------------------------------------------
def _reference_dir():
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline", "_ref")
    return d if os.path.isdir(os.path.join(d, "fsrl")) else None


def cpu_reference(sample_envs=32, cycles=1, threads=None):
    """CPU arm.  When `baseline/_ref` holds the reference (pip --no-deps --target install, built by
    __graft_entry__.build()), its OWN classes do the work that dominates the CPU time -- process_fn =
    BasePolicy.compute_gae_returns (numba gae_return) and PPOLagrangian.learn (eager torch autograd +
    Adam) run unmodified on the host cores; the third-party packages it imports but that are absent
    here (tianshou, gymnasium) are thin shims (oracle/refrun.py), and rollout collection -- which in
    the reference lives in tianshou's vector env / buffer and in pybullet -- is the oracle's in-process
    numpy env twin (no IPC: a lower bound on the reference's collect cost).  Without `baseline/_ref`
    the whole path is the oracle port."""
    ref_dir = _reference_dir()
    if ref_dir is not None:
        try:
            return cpu_reference_real(ref_dir, sample_envs, cycles, threads)
        except Exception as e:              # never lose the baseline: fall back to the port and say why
            print(f"[bench] reference classes unavailable ({type(e).__name__}: {e}); timing the oracle port",
                  file=sys.stderr)
    return cpu_reference_port(sample_envs, cycles, threads)


def cpu_reference_real(ref_dir, sample_envs, cycles, threads):
    import oracle.collector as ocol
    from oracle import refrun
    from oracle.envs import OracleVecEnv
    Batch = refrun.bootstrap(ref_dir)
    threads = threads or best_thread_count()
    torch.set_num_threads(threads)
    torch.manual_seed(SEED); np.random.seed(SEED)
    D, A, T = 8, 2, 300
    pol, actor, critics = refrun.ppo_lag_policy(D, A, HIDDEN, lr=5e-4, target_kl=float("inf"), max_grad_norm=0.5,
                                                cost_limit=10.0, gamma=0.99)      # ppol_cfg.py values
    act_fn = lambda obs: actor(obs)[0]                                           # (mu, sigma)
    env = OracleVecEnv(0, sample_envs, SEED); env.reset()
    buf = ocol.OracleBuffer(sample_envs * T, sample_envs, D, A)
    ctr = np.zeros(sample_envs, np.uint32)
    t0 = time.time()
    n = 0
    for _ in range(cycles):
        buf.reset()
        st = ocol.collect(env, act_fn, sample_envs, SEED, ctr, buf)
        pol.pre_update_fn(stats_train=st)                                        # PID step on the collect's cost
        idx = buf.sample_all()
        view = refrun.RingView(buf, Batch)
        batch = Batch(obs=torch.from_numpy(buf.obs[idx]), obs_next=torch.from_numpy(buf.obs_next[idx]),
                      act=torch.from_numpy(buf.act[idx]), rew=view.rew[idx], terminated=buf.terminated[idx],
                      truncated=buf.truncated[idx], info=Batch(cost=buf.cost[idx].astype(np.float64)))
        batch = pol.process_fn(batch, view, idx)                                 # ppo_lag.py:134-150
        pol.learn(batch, BATCH, REPEAT)                                          # ppo_lag.py:214-257
        n += st["n/st"]
    dt = time.time() - t0
    return {"value": n / dt, "unit": "env-steps/s", "cores": threads, "kind": "reference",
            "sample": f"{sample_envs} envs x {T} steps x {cycles} cycle(s) of c2 (2x256 MLP, batch 256, repeat 4): "
                      f"unmodified fsrl.policy.PPOLagrangian (process_fn + learn) from baseline/_ref with "
                      f"tianshou/gymnasium shims; collect = in-process numpy env twin (oracle), {dt:.1f} s"}


def cpu_reference_port(sample_envs=32, cycles=1, threads=None):
    """The reference's CPU path (FastCollector over per-env worker processes + numba GAE +
    eager-torch PPO update) cannot be installed here (tianshou/gymnasium/pybullet absent,
    no network): this times its restatement in oracle/ -- numpy env twin stepped in-process
    (no IPC: a lower bound on the reference's collect cost), C port of gae_return, torch-CPU
    autograd + Adam -- on a bounded sample of the c2 workload: `sample_envs` envs x 300 steps,
    2x256 MLP, batch 256, 4 repeats."""
    import oracle.collector as ocol
    from oracle import nets as onets, ppo as oppo
    from oracle.envs import OracleVecEnv
    threads = threads or best_thread_count()
    torch.set_num_threads(threads)
    torch.manual_seed(SEED); np.random.seed(SEED)
    D, A, T = 8, 2, 300
    actor = onets.GaussActor(D, A, list(HIDDEN))
    critics = [onets.ValueNet(D, list(HIDDEN)) for _ in range(2)]
    with torch.no_grad():
        actor.sigma_param.fill_(-0.5)
    for m in [actor] + critics:
        for l in m.modules():
            if isinstance(l, torch.nn.Linear):
                torch.nn.init.orthogonal_(l.weight); torch.nn.init.zeros_(l.bias)
    opt = torch.optim.Adam([p for m in [actor] + critics for p in m.parameters()], lr=5e-4)
    env = OracleVecEnv(0, sample_envs, SEED); env.reset()
    buf = ocol.OracleBuffer(sample_envs * T, sample_envs, D, A)
    ctr = np.zeros(sample_envs, np.uint32)
    t0 = time.time()
    n = 0
    for _ in range(cycles):
        buf.reset()
        st = ocol.collect(env, actor, sample_envs, SEED, ctr, buf)
        idx = buf.sample_all()
        b = {k: getattr(buf, k)[idx] for k in ("obs", "obs_next", "act", "rew", "cost", "terminated", "truncated")}
        b = oppo.process(actor, critics, b, 0.99, 0.95)
        oppo.learn(actor, critics, opt, b, BATCH, REPEAT, 0.0, max_grad_norm=0.5, target_kl=float("inf"))
        n += st["n/st"]
    dt = time.time() - t0
    return {"value": n / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": f"{sample_envs} envs x {T} steps x {cycles} cycle(s) of c2 (2x256 MLP, batch 256, "
                      f"repeat 4), in-process numpy env twin (no SubprocVectorEnv IPC), {dt:.1f} s"}


_BEST_THREADS = None


def best_thread_count():
    """The reference defaults to torch.set_num_threads(4) (ppol_cfg.py:11); tiny 2x256 MLPs do
    not scale with cores, so pick the fastest of {4, 8, 16, all} on a short calibration."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    lin = torch.nn.Sequential(torch.nn.Linear(8, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256),
                              torch.nn.ReLU(), torch.nn.Linear(256, 1))
    x = torch.randn(256, 8)
    best, best_t = 4, 1e9
    for th in sorted({4, 8, 16, os.cpu_count() or 4}):
        if th > (os.cpu_count() or 4):
            continue
        torch.set_num_threads(th)
        for _ in range(3):
            lin(x).sum().backward()
        t0 = time.time()
        for _ in range(30):
            lin(x).sum().backward()
        dt = time.time() - t0
        if dt < best_t:
            best, best_t = th, dt
    _BEST_THREADS = best
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    for _ in range(args.warmup):                      # untimed warm-up steps (thread pools, allocator) on a small sample
        cpu_reference(sample_envs=min(32, args.cpu_envs), cycles=1)
    t0 = time.time()
    vals = []
    for _ in range(args.steps):
        vals.append(cpu_reference(sample_envs=args.cpu_envs, cycles=1))
    dt = time.time() - t0
    n = args.steps * args.cpu_envs * 300
    v = n / dt
    cb = dict(vals[-1]); cb["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "env-steps/sec (collect+update) SafetyCarCircle-v0 PPO-Lag",
        "value": v, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2 shapes (2x256 MLP, batch 256, repeat 4) on a bounded sample of "
                               f"{args.cpu_envs} envs x 300 steps per step on the host cores; see "
                               "cpu_baseline.kind / sample for what ran (reference classes from baseline/_ref, "
                               "or the oracle port)"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cpu-envs", type=int, default=512,
                    help="envs of the bounded CPU sample (512 x 300 transitions ~ 10-15 s on 8 host cores)")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
