"""Synthetic hot-path inputs: seeded, env-major, ragged segments.  numpy only;
shared by tests/ and bench.py so both sides see the same bytes."""
from __future__ import annotations

import numpy as np


def synth_gae_inputs(n_env: int, T: int, seed: int = 10, p_term: float = 0.002, C: int = 2):
    """Returns dict of numpy arrays for a flat env-major buffer of N = n_env*T transitions:
    v, vnext (C,N) f32; rew, cost (N,) f32; terminated, truncated, unfinished (N,) bool."""
    rng = np.random.default_rng(seed)
    N = n_env * T
    term = rng.random(N) < p_term
    trunc = np.zeros(N, dtype=bool)
    if N:
        trunc[T - 1::T] = True
    term &= ~trunc
    unfinished = np.zeros(N, dtype=bool)
    v = rng.standard_normal((C, N)).astype(np.float32)
    vnext = rng.standard_normal((C, N)).astype(np.float32)
    rew = rng.normal(0.5, 1.0, N).astype(np.float32)
    cost = (rng.random(N) < 0.05).astype(np.float32)
    return dict(v=v, vnext=vnext, rew=rew, cost=cost, terminated=term, truncated=trunc,
                unfinished=unfinished)
