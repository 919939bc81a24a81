"""Generate tests/golden/* from the reference's OWN code (run in the build container only;
/root/reference does not exist on the GPU box, so the fixtures are committed).

  * gae_return / nstep_return are AST-extracted from
    /root/reference/fsrl/policy/base_policy.py:524-567 and compiled with the installed
    numba (the module itself cannot be imported: gymnasium/tianshou are absent).
  * LagrangianOptimizer is imported from /root/reference/fsrl/utils/optim_util.py.

Usage:  python -m oracle.make_golden
"""
from __future__ import annotations

import ast
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference_kernels():
    from numba import njit
    src = open(os.path.join(REF, "fsrl/policy/base_policy.py")).read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef)
           and n.name in ("gae_return", "nstep_return")]
    assert len(fns) == 2
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {"np": np, "njit": njit}
    exec(compile(mod, "base_policy_extract", "exec"), ns)
    return ns["gae_return"], ns["nstep_return"]


def load_reference_pid():
    spec = importlib.util.spec_from_file_location(
        "ref_optim_util", os.path.join(REF, "fsrl/utils/optim_util.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.LagrangianOptimizer


def synth_episodes(rng, n_env, T, p_term):
    """SURVEY.md 8(d) synthetic layout: env-major, truncated at each env's last step,
    Bernoulli(p_term) terminations that restart the episode."""
    N = n_env * T
    term = rng.random(N) < p_term
    trunc = np.zeros(N, dtype=bool)
    trunc[T - 1::T] = True
    term &= ~trunc
    return term, trunc


def main():
    os.makedirs(OUT, exist_ok=True)
    gae_ref, nstep_ref = load_reference_kernels()
    rng = np.random.default_rng(10)
    cases = {}
    # --- Appendix B hand cases -------------------------------------------------
    hand = [
        (np.array([1, 2, 3, 4], np.float32), np.array([2, 3, 4, 0], np.float32),
         np.array([1, 1, 1, 1], np.float64), np.array([0, 0, 0, 1], bool), 0.99, 0.95),
        (np.array([.5, .4, .3, .2, .1, 0, -.1], np.float32),
         np.array([.4, .3, 0, .1, 0, -.1, .7], np.float32),
         np.array([1, 0, 2, -1, .5, 0, 1], np.float64),
         np.array([0, 0, 1, 0, 0, 0, 1], bool), 0.99, 0.95),
        (np.array([0, 1], np.float32), np.array([0, 1], np.float32),
         np.array([0, 1], np.float64), np.array([False, True]), 0.1, 0.1),
    ]
    # --- seeded random cases, ragged segments, several sizes ---------------------
    for (n_env, T, p) in [(1, 1, 0.0), (1, 7, 0.3), (3, 5, 0.0), (4, 300, 0.002),
                          (16, 300, 0.01), (7, 513, 0.05), (33, 64, 0.2), (2, 4099, 0.001)]:
        N = n_env * T
        term, trunc = synth_episodes(rng, n_env, T, p)
        v = rng.standard_normal(N).astype(np.float32)
        vn = (rng.standard_normal(N).astype(np.float32) * ~term).astype(np.float32)
        r = rng.normal(0.5, 1.0, N)
        hand.append((v, vn, r, term | trunc, 0.99, 0.95))
    for k, (v, vn, r, e, g, l) in enumerate(hand):
        out = gae_ref(v, vn, r, e, g, l)
        assert out.dtype == np.float64
        cases[f"gae{k}_v"] = v
        cases[f"gae{k}_vn"] = vn
        cases[f"gae{k}_r"] = r
        cases[f"gae{k}_e"] = e
        cases[f"gae{k}_gl"] = np.array([g, l])
        cases[f"gae{k}_out"] = out
    cases["gae_count"] = np.array(len(hand))

    # --- nstep_return ------------------------------------------------------------
    ncases = [
        (np.array([1, 2, 3, 4, 5], np.float64), np.array([0, 0, 1, 0, 0], bool),
         np.array([[10], [20], [0], [40]], np.float32),
         np.array([[0, 1, 2, 3], [1, 2, 2, 4]], np.int64), 0.99, 2),
        (np.array([0, 1], np.float64), np.array([False, True]),
         np.array([[0], [1]], np.float32), np.array([[0, 1]], np.int64), 0.1, 1),
    ]
    for (buf, bsz, n_step) in [(50, 16, 1), (200, 64, 2), (1000, 256, 3), (333, 7, 5)]:
        metric = rng.normal(0.5, 1.0, buf)
        end = rng.random(buf) < 0.05
        start = rng.integers(0, buf, bsz)
        idx = [start]
        for _ in range(n_step - 1):
            last = idx[-1]
            nxt = np.where(end[last] | (last == buf - 1), last, last + 1)  # buffer.next()
            idx.append(nxt)
        idx = np.stack(idx).astype(np.int64)
        tq = rng.standard_normal((bsz, 1)).astype(np.float32)
        ncases.append((metric, end, tq, idx, 0.97, n_step))
    for k, (m, e, tq, idx, g, n) in enumerate(ncases):
        out = nstep_ref(m, e, tq, idx, g, n)
        cases[f"ns{k}_m"] = m
        cases[f"ns{k}_e"] = e
        cases[f"ns{k}_tq"] = tq
        cases[f"ns{k}_idx"] = idx
        cases[f"ns{k}_gn"] = np.array([g, n])
        cases[f"ns{k}_out"] = out
    cases["ns_count"] = np.array(len(ncases))
    np.savez_compressed(os.path.join(OUT, "returns_golden.npz"), **cases)

    # --- PID ------------------------------------------------------------------------
    Pid = load_reference_pid()
    pid_cases = []
    seqs = [((0.05, 0.0005, 0.1), 10.0, [25, 18, 12, 8, 9, 14]),
            ((0.05, 0.0005, 0.1), 25.0, list(np.round(rng.uniform(0, 60, 40), 3))),
            ((0.1, 0.01, 0.0), 5.0, list(np.round(rng.uniform(0, 12, 25), 3))),
            ((0.0, 0.035, 0.0), 10.0, list(np.round(rng.uniform(0, 30, 25), 3)))]
    for pid, limit, costs in seqs:
        o = Pid(pid)
        lam, integ, eold = [], [], []
        for c in costs:
            o.step(c, limit)
            lam.append(float(o.get_lag()))
            integ.append(float(o.error_integral))
            eold.append(float(o.error_old))
        pid_cases.append({"pid": list(pid), "limit": limit, "costs": [float(c) for c in costs],
                          "lagrangian": lam, "error_integral": integ, "error_old": eold,
                          "state_dict": {k: (list(v) if isinstance(v, tuple) else float(v))
                                         for k, v in o.state_dict().items()}})
    with open(os.path.join(OUT, "pid_golden.json"), "w") as f:
        json.dump(pid_cases, f, indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
