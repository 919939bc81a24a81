"""Oracle (test infrastructure): CPU twin of the batched device environments in
fsrl_b200/csrc/envs.cuh, vectorised over envs in numpy float32.

The reference's environments (bullet_safety_gym -> pybullet, safety_gymnasium -> mujoco;
called at /root/reference/fsrl/data/fast_collector.py:286) are absent and irreproducible
(SURVEY.md F5): env parity with the reference is UNPINNED.  These are our documented models;
every op is IEEE-exact (+ - * / sqrt), written in the same order as the CUDA code, so device
trajectories match this twin bit-for-bit given identical actions.
"""
from __future__ import annotations

import numpy as np

from .philox import KEY_GOAL, KEY_RESET, philox4x32, usym

f32 = np.float32
KINDS = {"car_circle": 0, "car_run": 1, "ball_circle": 2, "ball_run": 3, "ant_circle": 4,
         "point_goal": 5}
DIMS = {0: (8, 2, 6, 300), 1: (7, 2, 7, 200), 2: (8, 2, 4, 200), 3: (7, 2, 5, 100),
        4: (34, 8, 30, 500), 5: (60, 2, 28, 1000)}   # D, A, S, T

LIDAR_EDGE_C = np.array([1.0, 0.92387953, 0.70710678, 0.38268343, 0.0, -0.38268343, -0.70710678,
                         -0.92387953, -1.0, -0.92387953, -0.70710678, -0.38268343, 0.0,
                         0.38268343, 0.70710678, 0.92387953], dtype=f32)
LIDAR_EDGE_S = np.array([0.0, 0.38268343, 0.70710678, 0.92387953, 1.0, 0.92387953, 0.70710678,
                         0.38268343, 0.0, -0.38268343, -0.70710678, -0.92387953, -1.0,
                         -0.92387953, -0.70710678, -0.38268343], dtype=f32)


def _rotate(c, s, d):
    d2 = d * d
    sn = d * (f32(1) - (d2 / f32(6)) * (f32(1) - d2 / f32(20)))
    cs = f32(1) - (d2 / f32(2)) * (f32(1) - (d2 / f32(12)) * (f32(1) - d2 / f32(30)))
    c2 = c * cs - s * sn
    s2 = s * cs + c * sn
    n = np.sqrt(c2 * c2 + s2 * s2)
    return c2 / n, s2 / n


def _heading_from_box(a, b):
    n2 = a * a + b * b
    small = n2 < f32(1e-12)
    n = np.sqrt(np.where(small, f32(1), n2))
    c = np.where(small, f32(1), a / n).astype(f32)
    s = np.where(small, f32(0), b / n).astype(f32)
    return c, s


def _car_advance(st, a0, a1, vmax, wmax, av, aw, dt):
    x, y, c, s, v, w = (st[i] for i in range(6))
    v = v + (a0 * f32(vmax) - v) * f32(av)
    w = w + (a1 * f32(wmax) - w) * f32(aw)
    c, s = _rotate(c, s, w * f32(dt))
    x = x + (v * c) * f32(dt)
    y = y + (v * s) * f32(dt)
    st[0], st[1], st[2], st[3], st[4], st[5] = x, y, c, s, v, w


class OracleVecEnv:
    """E independent envs, state SoA (S, E) float32.  API: reset(ids), observe(), step(act)."""

    def __init__(self, kind, n_env, seed):
        self.kind = KINDS[kind] if isinstance(kind, str) else int(kind)
        self.D, self.A, self.S, self.T = DIMS[self.kind]
        self.E = n_env
        self.seed = np.uint32(seed)
        self.st = np.zeros((self.S, n_env), dtype=f32)
        self.ep_idx = np.zeros(n_env, dtype=np.uint32)   # episodes started per env
        self.t = np.zeros(n_env, dtype=np.int32)

    # ---- reset ------------------------------------------------------------------------------
    def reset(self, ids=None):
        ids = np.arange(self.E) if ids is None else np.asarray(ids)
        env = ids.astype(np.uint32)
        ep = self.ep_idx[ids]
        r = philox4x32(env, ep, 0, 0, self.seed, KEY_RESET)
        st = np.zeros((self.S, len(ids)), dtype=f32)
        k = self.kind
        if k == 0:
            st[0] = usym(r[0]) * f32(0.3); st[1] = usym(r[1]) * f32(0.3)
            st[2], st[3] = _heading_from_box(usym(r[2]), usym(r[3]))
        elif k == 1:
            st[1] = usym(r[0]) * f32(0.2)
            st[2], st[3] = _heading_from_box(np.ones(len(ids), f32), usym(r[1]) * f32(0.3))
        elif k == 2:
            st[0] = usym(r[0]) * f32(0.3); st[1] = usym(r[1]) * f32(0.3)
        elif k == 3:
            st[1] = usym(r[0]) * f32(0.2)
        elif k == 4:
            st[0] = usym(r[0]) * f32(0.5); st[1] = usym(r[1]) * f32(0.5)
            st[2], st[3] = _heading_from_box(usym(r[2]), usym(r[3]))
            q = philox4x32(env, ep, 1, 0, self.seed, KEY_RESET)
            q2 = philox4x32(env, ep, 2, 0, self.seed, KEY_RESET)
            for j in range(4):
                st[6 + j] = usym(q[j]) * f32(0.1)
                st[10 + j] = usym(q2[j]) * f32(0.1)
        elif k == 5:
            st[0] = usym(r[0]) * f32(0.5); st[1] = usym(r[1]) * f32(0.5)
            st[2], st[3] = _heading_from_box(usym(r[2]), usym(r[3]))
            g = philox4x32(env, ep, 0, 0, self.seed, KEY_GOAL)
            st[6] = usym(g[0]) * f32(2.0); st[7] = usym(g[1]) * f32(2.0)
            for h in range(5):
                q = philox4x32(env, ep, 1 + h, 0, self.seed, KEY_RESET)
                if h < 4:
                    for j in range(4):
                        st[9 + 4 * h + j] = usym(q[j]) * f32(2.0)
                else:
                    st[25] = usym(q[0]) * f32(2.0); st[26] = usym(q[1]) * f32(2.0)
        self.st[:, ids] = st
        self.ep_idx[ids] += np.uint32(1)
        self.t[ids] = 0
        return self.observe(ids)

    # ---- observation ---------------------------------------------------------------------------
    def observe(self, ids=None):
        st = self.st if ids is None else self.st[:, ids]
        n = st.shape[1]
        o = np.zeros((n, self.D), dtype=f32)
        k = self.kind
        if k in (0, 4):
            R, WMAX = (f32(1.5), f32(3.0)) if k == 0 else (f32(3.0), f32(2.0))
            x, y, c, s, v, w = (st[i] for i in range(6))
            r = np.sqrt(x * x + y * y)
            o[:, 0] = x / R; o[:, 1] = y / R; o[:, 2] = v * c; o[:, 3] = v * s
            o[:, 4] = c; o[:, 5] = s; o[:, 6] = w / WMAX; o[:, 7] = (r - R) / R
            if k == 4:
                aq = np.zeros(n, f32)
                for j in range(8):
                    o[:, 8 + j] = st[6 + j]
                    o[:, 16 + j] = st[14 + j] * f32(0.1)
                    o[:, 24 + j] = st[22 + j]
                    aq = aq + np.abs(st[6 + j])
                o[:, 32] = v / f32(2.0)
                o[:, 33] = f32(0.5) + aq * f32(0.0125)
        elif k == 1:
            o[:, 0] = st[1]; o[:, 1] = st[4] * st[2]; o[:, 2] = st[4] * st[3]; o[:, 3] = st[2]
            o[:, 4] = st[3]; o[:, 5] = st[5] / f32(3.0); o[:, 6] = st[4] / f32(1.2)
        elif k == 2:
            R = f32(1.5)
            x, y, vx, vy = (st[i] for i in range(4))
            r = np.sqrt(x * x + y * y)
            rg = r + f32(1e-6)
            o[:, 0] = x / R; o[:, 1] = y / R; o[:, 2] = vx; o[:, 3] = vy; o[:, 4] = (r - R) / R
            o[:, 5] = np.sqrt(vx * vx + vy * vy); o[:, 6] = x / rg; o[:, 7] = y / rg
        elif k == 3:
            y, vx, vy = st[1], st[2], st[3]
            sp = np.sqrt(vx * vx + vy * vy)
            o[:, 0] = y; o[:, 1] = vx; o[:, 2] = vy; o[:, 3] = sp; o[:, 4] = sp - f32(1.5)
            o[:, 5] = np.abs(y) - f32(0.6); o[:, 6] = st[0] / f32(10.0)
        elif k == 5:
            x, y, c, s, v, w = (st[i] for i in range(6))
            o[:, 0] = (v - st[27]) / f32(0.05); o[:, 1] = v * w; o[:, 2] = f32(9.81)
            o[:, 3] = v; o[:, 8] = w; o[:, 9] = c; o[:, 10] = f32(0.0) - s

            def lidar(col0, ox, oy):
                dx = ox - x; dy = oy - y
                rx = c * dx + s * dy
                ry = c * dy - s * dx
                d = np.sqrt(rx * rx + ry * ry)
                val = np.maximum(f32(0), f32(1) - d / f32(3.0))
                b = np.zeros(n, dtype=np.int64)
                for kk in range(16):
                    k1 = (kk + 1) & 15
                    c0 = LIDAR_EDGE_C[kk] * ry - LIDAR_EDGE_S[kk] * rx
                    c1 = LIDAR_EDGE_C[k1] * ry - LIDAR_EDGE_S[k1] * rx
                    b = np.where((c0 >= 0) & (c1 < 0), kk, b)
                idx = np.arange(n)
                o[idx, col0 + b] = np.maximum(o[idx, col0 + b], val)

            lidar(12, st[6], st[7])
            for h in range(8):
                lidar(28, st[9 + 2 * h], st[10 + 2 * h])
            lidar(44, st[25], st[26])
        return o

    # ---- step ----------------------------------------------------------------------------------
    def step(self, act, ids=None):
        """act: (n, A) float32 already mapped to the env's action range.  Returns
        (obs_next, rew, cost, terminated, truncated) for the stepped envs; does NOT reset."""
        ids = np.arange(self.E) if ids is None else np.asarray(ids)
        act = np.asarray(act, dtype=f32)
        st = [self.st[i, ids].copy() for i in range(self.S)]
        k = self.kind
        n = len(ids)
        term = np.zeros(n, dtype=bool)
        if k == 0:
            _car_advance(st, act[:, 0], act[:, 1], 1.5, 3.0, 0.2, 0.3, 0.05)
            x, y = st[0], st[1]
            vx, vy = st[4] * st[2], st[4] * st[3]
            r = np.sqrt(x * x + y * y)
            rew = (x * vy - y * vx) / (f32(1.5) * (f32(1) + np.abs(r - f32(1.5))))
            cost = (np.abs(x) > f32(1.125)).astype(f32)
        elif k == 1:
            x_old = st[0].copy()
            _car_advance(st, act[:, 0], act[:, 1], 1.5, 3.0, 0.2, 0.3, 0.05)
            rew = ((st[0] - x_old) / f32(0.05)) * f32(2.0)
            cost = ((np.abs(st[1]) > f32(0.6)) | (st[4] > f32(1.2))).astype(f32)
            st[6] = st[6] + cost
        elif k in (2, 3):
            x_old = st[0].copy()
            st[2] = st[2] + (act[:, 0] * f32(4.0) - f32(2.0) * st[2]) * f32(0.05)
            st[3] = st[3] + (act[:, 1] * f32(4.0) - f32(2.0) * st[3]) * f32(0.05)
            st[0] = st[0] + st[2] * f32(0.05)
            st[1] = st[1] + st[3] * f32(0.05)
            x, y, vx, vy = st[0], st[1], st[2], st[3]
            if k == 2:
                r = np.sqrt(x * x + y * y)
                rew = (x * vy - y * vx) / (f32(1.5) * (f32(1) + np.abs(r - f32(1.5))))
                cost = (np.abs(x) > f32(1.125)).astype(f32)
            else:
                sp = np.sqrt(vx * vx + vy * vy)
                rew = ((st[0] - x_old) / f32(0.05)) * f32(2.5)
                cost = ((np.abs(y) > f32(0.6)) | (sp > f32(1.5))).astype(f32)
                st[4] = st[4] + cost
        elif k == 4:
            thrust = np.zeros(n, f32); turn = np.zeros(n, f32); ctrl = np.zeros(n, f32)
            for j in range(8):
                q, qd, a = st[6 + j], st[14 + j], act[:, j]
                qd = qd + (((f32(20.0) * a) - (f32(10.0) * q)) - (f32(4.0) * qd)) * f32(0.05)
                q = q + qd * f32(0.05)
                st[6 + j], st[14 + j], st[22 + j] = q, qd, a
                if j < 4:
                    thrust = thrust + q
                else:
                    turn = turn + q
                ctrl = ctrl + a * a
            f = np.minimum(f32(1), np.maximum(f32(-1), thrust * f32(0.25)))
            g = np.minimum(f32(1), np.maximum(f32(-1), turn * f32(0.25)))
            _car_advance(st, f, g, 2.0, 2.0, 0.1, 0.15, 0.05)
            x, y = st[0], st[1]
            vx, vy = st[4] * st[2], st[4] * st[3]
            r = np.sqrt(x * x + y * y)
            rew = (x * vy - y * vx) / (f32(3.0) * (f32(1) + np.abs(r - f32(3.0)))) - f32(0.005) * ctrl
            cost = (np.abs(x) > f32(2.25)).astype(f32)
        elif k == 5:
            dxo = st[6] - st[0]; dyo = st[7] - st[1]
            dist_old = np.sqrt(dxo * dxo + dyo * dyo)
            st[27] = st[4].copy()
            _car_advance(st, act[:, 0], act[:, 1], 1.0, 3.0, 0.2, 0.3, 0.05)
            st[0] = np.minimum(f32(2), np.maximum(f32(-2), st[0]))
            st[1] = np.minimum(f32(2), np.maximum(f32(-2), st[1]))
            dxn = st[6] - st[0]; dyn = st[7] - st[1]
            dist = np.sqrt(dxn * dxn + dyn * dyn)
            rew = dist_old - dist
            hit = dist <= f32(0.3)
            rew = np.where(hit, rew + f32(1), rew).astype(f32)
            st[8] = np.where(hit, st[8] + f32(1), st[8]).astype(f32)
            if hit.any():
                env = ids.astype(np.uint32)
                ep = (self.ep_idx[ids] - np.uint32(1)).astype(np.uint32)
                g = philox4x32(env, ep, (np.uint32(16) + st[8].astype(np.uint32)), 0, self.seed, KEY_GOAL)
                st[6] = np.where(hit, usym(g[0]) * f32(2.0), st[6]).astype(f32)
                st[7] = np.where(hit, usym(g[1]) * f32(2.0), st[7]).astype(f32)
            cost = np.zeros(n, f32)
            for h in range(8):
                dx = st[9 + 2 * h] - st[0]; dy = st[10 + 2 * h] - st[1]
                cost = np.where(dx * dx + dy * dy <= f32(0.2) * f32(0.2), f32(1), cost).astype(f32)
        for i in range(self.S):
            self.st[i, ids] = st[i]
        self.t[ids] += 1
        trunc = self.t[ids] >= self.T
        return self.observe(ids), rew.astype(f32), cost.astype(f32), term, trunc
