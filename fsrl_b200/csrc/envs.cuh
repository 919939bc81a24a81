// Batched on-device safe-RL environments (the "env.step" the reference delegates to
// pybullet / mujoco worker processes: fsrl/data/fast_collector.py:286, tianshou vector envs).
//
// The reference's physics engines are absent and irreproducible, so these are
// OUR documented analytic models with the task structure of Bullet-Safety-Gym's Circle / Run
// tasks (dense reward, binary cost, fixed horizon, truncation only).  Every arithmetic step
// uses only IEEE-exact operations (+ - * / sqrt, no FMA contraction, polynomial sin/cos), so
// the CPU twin in oracle/envs.py reproduces trajectories BIT-EXACTLY from the same actions.
//
// State lives in registers of the thread that owns the env; SoA [S][E] in HBM between steps.
#pragma once
#include "common.cuh"

namespace fsrl {

enum EnvKind { ENV_CAR_CIRCLE = 0, ENV_CAR_RUN = 1, ENV_BALL_CIRCLE = 2, ENV_BALL_RUN = 3,
               ENV_ANT_CIRCLE = 4, ENV_POINT_GOAL = 5, ENV_KIND_COUNT = 6 };

constexpr int ENV_MAX_D = 64;
constexpr int ENV_MAX_A = 8;
constexpr int ENV_MAX_S = 32;

// exact-op helpers (never contracted into FMA)
__device__ __forceinline__ float xm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xq(float a) { return __fsqrt_rn(a); }

// rotate the unit heading (c, s) by a small angle d with polynomial sin/cos, renormalise
__device__ __forceinline__ void rotate_heading(float& c, float& s, float d) {
    const float d2 = xm(d, d);
    // sn = d * (1 - d2/6 * (1 - d2/20));  cs = 1 - d2/2 * (1 - d2/12 * (1 - d2/30))
    const float sn = xm(d, xs(1.0f, xm(xd(d2, 6.0f), xs(1.0f, xd(d2, 20.0f)))));
    const float cs = xs(1.0f, xm(xd(d2, 2.0f), xs(1.0f, xm(xd(d2, 12.0f), xs(1.0f, xd(d2, 30.0f))))));
    const float c2 = xs(xm(c, cs), xm(s, sn));
    const float s2 = xa(xm(s, cs), xm(c, sn));
    const float n = xq(xa(xm(c2, c2), xm(s2, s2)));
    c = xd(c2, n);
    s = xd(s2, n);
}

// uniform in [-1, 1): 2*u - 1 with u = (x >> 8) * 2^-24  (exact)
__device__ __forceinline__ float usym(uint32_t x) {
    return xs(xm((float)(x >> 8), 2.0f / 16777216.0f), 1.0f);
}

constexpr uint32_t KEY_RESET = 0x52534554u;  // 'RSET'
constexpr uint32_t KEY_ACT = 0x4143544Eu;    // 'ACTN'
constexpr uint32_t KEY_GOAL = 0x474F414Cu;   // 'GOAL'

// ---- model constants (mirrored in oracle/envs.py) -------------------------------------------
namespace carc {
constexpr float DT = 0.05f, R = 1.5f, XLIM = 1.125f, VMAX = 1.5f, WMAX = 3.0f, AV = 0.2f, AW = 0.3f;
}
namespace carr {
constexpr float DT = 0.05f, YLIM = 0.6f, VLIM = 1.2f, VMAX = 1.5f, WMAX = 3.0f, AV = 0.2f, AW = 0.3f, RSCALE = 2.0f;
}
namespace ball {
constexpr float DT = 0.05f, R = 1.5f, XLIM = 1.125f, ACC = 4.0f, DRAG = 2.0f, YLIM = 0.6f, VLIM = 1.5f, RSCALE = 2.5f;
}
namespace ant {
constexpr float DT = 0.05f, R = 3.0f, XLIM = 2.25f, VMAX = 2.0f, WMAX = 2.0f, AV = 0.1f, AW = 0.15f;
constexpr float KA = 20.0f, KQ = 10.0f, KD = 4.0f;
}
namespace pgoal {
constexpr float DT = 0.05f, VMAX = 1.0f, WMAX = 3.0f, AV = 0.2f, AW = 0.3f, ARENA = 2.0f;
constexpr float GOAL_R = 0.3f, HAZ_R = 0.2f, LIDAR_MAX = 3.0f;
constexpr int NHAZ = 8, NBIN = 16;
}

struct EnvDims { int D, A, S, T; };

__host__ __device__ inline EnvDims env_dims(int kind) {
    switch (kind) {
        case ENV_CAR_CIRCLE: return {8, 2, 6, 300};
        case ENV_CAR_RUN: return {7, 2, 7, 200};
        case ENV_BALL_CIRCLE: return {8, 2, 4, 200};
        case ENV_BALL_RUN: return {7, 2, 5, 100};
        case ENV_ANT_CIRCLE: return {34, 8, 30, 500};
        case ENV_POINT_GOAL: return {60, 2, 28, 1000};
        default: return {0, 0, 0, 0};
    }
}

// ---------------------------------------------------------------------------------------------
// Car (unicycle with first-order actuator lag).  state: x, y, c, s, v, w [, x0 (run)]
// ---------------------------------------------------------------------------------------------
template <int KIND>
struct Env;

__device__ __forceinline__ void heading_from_box(float a, float b, float& c, float& s) {
    // direction of a uniform point of the square (documented: not a uniform angle)
    float n2 = xa(xm(a, a), xm(b, b));
    if (n2 < 1e-12f) { c = 1.0f; s = 0.0f; return; }
    const float n = xq(n2);
    c = xd(a, n);
    s = xd(b, n);
}

__device__ __forceinline__ void car_advance(float* st, float a0, float a1, float vmax, float wmax,
                                            float av, float aw, float dt) {
    float x = st[0], y = st[1], c = st[2], s = st[3], v = st[4], w = st[5];
    v = xa(v, xm(xs(xm(a0, vmax), v), av));
    w = xa(w, xm(xs(xm(a1, wmax), w), aw));
    rotate_heading(c, s, xm(w, dt));
    x = xa(x, xm(xm(v, c), dt));
    y = xa(y, xm(xm(v, s), dt));
    st[0] = x; st[1] = y; st[2] = c; st[3] = s; st[4] = v; st[5] = w;
}

template <>
struct Env<ENV_CAR_CIRCLE> {
    static constexpr int D = 8, A = 2, S = 6, T = 300;
    __device__ static void reset(float* st, uint32_t seed, uint32_t env, uint32_t ep) {
        uint32_t r[4];
        Philox::gen(env, ep, 0u, 0u, seed, KEY_RESET, r);
        st[0] = xm(usym(r[0]), 0.3f);
        st[1] = xm(usym(r[1]), 0.3f);
        heading_from_box(usym(r[2]), usym(r[3]), st[2], st[3]);
        st[4] = 0.0f; st[5] = 0.0f;
    }
    __device__ static void observe(const float* st, float* o) {
        using namespace carc;
        const float x = st[0], y = st[1], c = st[2], s = st[3], v = st[4], w = st[5];
        const float r = xq(xa(xm(x, x), xm(y, y)));
        o[0] = xd(x, R); o[1] = xd(y, R); o[2] = xm(v, c); o[3] = xm(v, s);
        o[4] = c; o[5] = s; o[6] = xd(w, WMAX); o[7] = xd(xs(r, R), R);
    }
    __device__ static void step(float* st, const float* a, uint32_t, uint32_t, uint32_t,
                                float& rew, float& cost, bool& term) {
        using namespace carc;
        car_advance(st, a[0], a[1], VMAX, WMAX, AV, AW, DT);
        const float x = st[0], y = st[1], vx = xm(st[4], st[2]), vy = xm(st[4], st[3]);
        const float r = xq(xa(xm(x, x), xm(y, y)));
        // reward = (x*vy - y*vx) / (R * (1 + |r - R|))
        rew = xd(xs(xm(x, vy), xm(y, vx)), xm(R, xa(1.0f, fabsf(xs(r, R)))));
        cost = (fabsf(x) > XLIM) ? 1.0f : 0.0f;
        term = false;
    }
};

template <>
struct Env<ENV_CAR_RUN> {
    static constexpr int D = 7, A = 2, S = 7, T = 200;
    __device__ static void reset(float* st, uint32_t seed, uint32_t env, uint32_t ep) {
        uint32_t r[4];
        Philox::gen(env, ep, 0u, 0u, seed, KEY_RESET, r);
        st[0] = 0.0f;
        st[1] = xm(usym(r[0]), 0.2f);
        heading_from_box(1.0f, xm(usym(r[1]), 0.3f), st[2], st[3]);
        st[4] = 0.0f; st[5] = 0.0f; st[6] = 0.0f;
    }
    __device__ static void observe(const float* st, float* o) {
        using namespace carr;
        o[0] = st[1]; o[1] = xm(st[4], st[2]); o[2] = xm(st[4], st[3]); o[3] = st[2]; o[4] = st[3];
        o[5] = xd(st[5], WMAX); o[6] = xd(st[4], VLIM);
    }
    __device__ static void step(float* st, const float* a, uint32_t, uint32_t, uint32_t,
                                float& rew, float& cost, bool& term) {
        using namespace carr;
        const float x_old = st[0];
        car_advance(st, a[0], a[1], VMAX, WMAX, AV, AW, DT);
        rew = xm(xd(xs(st[0], x_old), DT), RSCALE);
        cost = (fabsf(st[1]) > YLIM || st[4] > VLIM) ? 1.0f : 0.0f;
        st[6] = xa(st[6], cost);
        term = false;
    }
};

// ---------------------------------------------------------------------------------------------
// Ball (force-controlled point mass with linear drag).  state: x, y, vx, vy [, x0]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ball_advance(float* st, float a0, float a1) {
    using namespace ball;
    float x = st[0], y = st[1], vx = st[2], vy = st[3];
    vx = xa(vx, xm(xs(xm(a0, ACC), xm(DRAG, vx)), DT));
    vy = xa(vy, xm(xs(xm(a1, ACC), xm(DRAG, vy)), DT));
    x = xa(x, xm(vx, DT));
    y = xa(y, xm(vy, DT));
    st[0] = x; st[1] = y; st[2] = vx; st[3] = vy;
}

template <>
struct Env<ENV_BALL_CIRCLE> {
    static constexpr int D = 8, A = 2, S = 4, T = 200;
    __device__ static void reset(float* st, uint32_t seed, uint32_t env, uint32_t ep) {
        uint32_t r[4];
        Philox::gen(env, ep, 0u, 0u, seed, KEY_RESET, r);
        st[0] = xm(usym(r[0]), 0.3f); st[1] = xm(usym(r[1]), 0.3f); st[2] = 0.0f; st[3] = 0.0f;
    }
    __device__ static void observe(const float* st, float* o) {
        using namespace ball;
        const float x = st[0], y = st[1], vx = st[2], vy = st[3];
        const float r = xq(xa(xm(x, x), xm(y, y)));
        const float rg = xa(r, 1e-6f);
        o[0] = xd(x, R); o[1] = xd(y, R); o[2] = vx; o[3] = vy; o[4] = xd(xs(r, R), R);
        o[5] = xq(xa(xm(vx, vx), xm(vy, vy))); o[6] = xd(x, rg); o[7] = xd(y, rg);
    }
    __device__ static void step(float* st, const float* a, uint32_t, uint32_t, uint32_t,
                                float& rew, float& cost, bool& term) {
        using namespace ball;
        ball_advance(st, a[0], a[1]);
        const float x = st[0], y = st[1], vx = st[2], vy = st[3];
        const float r = xq(xa(xm(x, x), xm(y, y)));
        rew = xd(xs(xm(x, vy), xm(y, vx)), xm(R, xa(1.0f, fabsf(xs(r, R)))));
        cost = (fabsf(x) > XLIM) ? 1.0f : 0.0f;
        term = false;
    }
};

template <>
struct Env<ENV_BALL_RUN> {
    static constexpr int D = 7, A = 2, S = 5, T = 100;
    __device__ static void reset(float* st, uint32_t seed, uint32_t env, uint32_t ep) {
        uint32_t r[4];
        Philox::gen(env, ep, 0u, 0u, seed, KEY_RESET, r);
        st[0] = 0.0f; st[1] = xm(usym(r[0]), 0.2f); st[2] = 0.0f; st[3] = 0.0f; st[4] = 0.0f;
    }
    __device__ static void observe(const float* st, float* o) {
        using namespace ball;
        const float y = st[1], vx = st[2], vy = st[3];
        const float sp = xq(xa(xm(vx, vx), xm(vy, vy)));
        o[0] = y; o[1] = vx; o[2] = vy; o[3] = sp; o[4] = xs(sp, VLIM); o[5] = xs(fabsf(y), YLIM);
        o[6] = xd(st[0], 10.0f);
    }
    __device__ static void step(float* st, const float* a, uint32_t, uint32_t, uint32_t,
                                float& rew, float& cost, bool& term) {
        using namespace ball;
        const float x_old = st[0];
        ball_advance(st, a[0], a[1]);
        const float vx = st[2], vy = st[3];
        const float sp = xq(xa(xm(vx, vx), xm(vy, vy)));
        rew = xm(xd(xs(st[0], x_old), DT), RSCALE);
        cost = (fabsf(st[1]) > YLIM || sp > VLIM) ? 1.0f : 0.0f;
        st[4] = xa(st[4], cost);
        term = false;
    }
};

// ---------------------------------------------------------------------------------------------
// Ant-Circle (D = 34, A = 8): a torso that moves like the car, driven by 8 actuated joints
// modelled as damped oscillators.  Joints 0-3 contribute thrust, 4-7 contribute turning.
// state: x, y, c, s, v, w, q[8], qd[8], a_prev[8]
// ---------------------------------------------------------------------------------------------
template <>
struct Env<ENV_ANT_CIRCLE> {
    static constexpr int D = 34, A = 8, S = 30, T = 500;
    __device__ static void reset(float* st, uint32_t seed, uint32_t env, uint32_t ep) {
        uint32_t r[4];
        Philox::gen(env, ep, 0u, 0u, seed, KEY_RESET, r);
        st[0] = xm(usym(r[0]), 0.5f);
        st[1] = xm(usym(r[1]), 0.5f);
        heading_from_box(usym(r[2]), usym(r[3]), st[2], st[3]);
        st[4] = 0.0f; st[5] = 0.0f;
        uint32_t q[4];
        Philox::gen(env, ep, 1u, 0u, seed, KEY_RESET, q);
        uint32_t q2[4];
        Philox::gen(env, ep, 2u, 0u, seed, KEY_RESET, q2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st[6 + j] = xm(usym(q[j]), 0.1f);
            st[10 + j] = xm(usym(q2[j]), 0.1f);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) st[14 + j] = 0.0f;
    }
    __device__ static void observe(const float* st, float* o) {
        using namespace ant;
        const float x = st[0], y = st[1], c = st[2], s = st[3], v = st[4], w = st[5];
        const float r = xq(xa(xm(x, x), xm(y, y)));
        o[0] = xd(x, R); o[1] = xd(y, R); o[2] = xm(v, c); o[3] = xm(v, s);
        o[4] = c; o[5] = s; o[6] = xd(w, WMAX); o[7] = xd(xs(r, R), R);
        float aq = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[8 + j] = st[6 + j];
            o[16 + j] = xm(st[14 + j], 0.1f);
            o[24 + j] = st[22 + j];
            aq = xa(aq, fabsf(st[6 + j]));
        }
        o[32] = xd(v, VMAX);
        o[33] = xa(0.5f, xm(aq, 0.0125f));
    }
    __device__ static void step(float* st, const float* a, uint32_t, uint32_t, uint32_t,
                                float& rew, float& cost, bool& term) {
        using namespace ant;
        float thrust = 0.0f, turn = 0.0f, ctrl = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float q = st[6 + j], qd = st[14 + j];
            // qd += (KA*a - KQ*q - KD*qd) * dt ; q += qd * dt
            qd = xa(qd, xm(xs(xs(xm(KA, a[j]), xm(KQ, q)), xm(KD, qd)), DT));
            q = xa(q, xm(qd, DT));
            st[6 + j] = q; st[14 + j] = qd; st[22 + j] = a[j];
            if (j < 4) thrust = xa(thrust, q); else turn = xa(turn, q);
            ctrl = xa(ctrl, xm(a[j], a[j]));
        }
        // joint deflection (bounded to [-1,1]) commands the torso
        float f = fminf(1.0f, fmaxf(-1.0f, xm(thrust, 0.25f)));
        float g = fminf(1.0f, fmaxf(-1.0f, xm(turn, 0.25f)));
        car_advance(st, f, g, VMAX, WMAX, AV, AW, DT);
        const float x = st[0], y = st[1], vx = xm(st[4], st[2]), vy = xm(st[4], st[3]);
        const float r = xq(xa(xm(x, x), xm(y, y)));
        rew = xs(xd(xs(xm(x, vy), xm(y, vx)), xm(R, xa(1.0f, fabsf(xs(r, R))))), xm(0.005f, ctrl));
        cost = (fabsf(x) > XLIM) ? 1.0f : 0.0f;
        term = false;
    }
};

// ---------------------------------------------------------------------------------------------
// Point-Goal1 (D = 60, A = 2, T = 1000): unicycle robot, one goal (re-sampled when reached),
// 8 hazards, 1 vase; three 16-bin pseudo-lidars computed with exact ops (sector membership by
// cross products against constant bin-edge directions, no atan2).
// state: x, y, c, s, v, w, gx, gy, goal_count, haz[8][2], vase[2], v_prev, w_prev
// ---------------------------------------------------------------------------------------------
__device__ __constant__ float LIDAR_EDGE_C[16] = {
    1.0f, 0.92387953f, 0.70710678f, 0.38268343f, 0.0f, -0.38268343f, -0.70710678f, -0.92387953f,
    -1.0f, -0.92387953f, -0.70710678f, -0.38268343f, 0.0f, 0.38268343f, 0.70710678f, 0.92387953f};
__device__ __constant__ float LIDAR_EDGE_S[16] = {
    0.0f, 0.38268343f, 0.70710678f, 0.92387953f, 1.0f, 0.92387953f, 0.70710678f, 0.38268343f,
    0.0f, -0.38268343f, -0.70710678f, -0.92387953f, -1.0f, -0.92387953f, -0.70710678f, -0.38268343f};

__device__ __forceinline__ void lidar_add(float* bins, float rx, float ry) {
    // rx, ry: object position in the robot frame.  Writes max(closeness) into its sector.
    using namespace pgoal;
    const float d = xq(xa(xm(rx, rx), xm(ry, ry)));
    const float val = fmaxf(0.0f, xs(1.0f, xd(d, LIDAR_MAX)));
    int bin = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int k1 = (k + 1) & 15;
        const float c0 = xs(xm(LIDAR_EDGE_C[k], ry), xm(LIDAR_EDGE_S[k], rx));     // cross(edge_k, r)
        const float c1 = xs(xm(LIDAR_EDGE_C[k1], ry), xm(LIDAR_EDGE_S[k1], rx));   // cross(edge_k+1, r)
        if (c0 >= 0.0f && c1 < 0.0f) bin = k;
    }
    bins[bin] = fmaxf(bins[bin], val);
}

template <>
struct Env<ENV_POINT_GOAL> {
    static constexpr int D = 60, A = 2, S = 28, T = 1000;
    __device__ static void sample_goal(float* st, uint32_t seed, uint32_t env, uint32_t ep, uint32_t k) {
        uint32_t r[4];
        Philox::gen(env, ep, k, 0u, seed, KEY_GOAL, r);
        st[6] = xm(usym(r[0]), pgoal::ARENA);
        st[7] = xm(usym(r[1]), pgoal::ARENA);
    }
    __device__ static void reset(float* st, uint32_t seed, uint32_t env, uint32_t ep) {
        using namespace pgoal;
        uint32_t r[4];
        Philox::gen(env, ep, 0u, 0u, seed, KEY_RESET, r);
        st[0] = xm(usym(r[0]), 0.5f);
        st[1] = xm(usym(r[1]), 0.5f);
        heading_from_box(usym(r[2]), usym(r[3]), st[2], st[3]);
        st[4] = 0.0f; st[5] = 0.0f;
        sample_goal(st, seed, env, ep, 0u);
        st[8] = 0.0f;
#pragma unroll
        for (int h = 0; h < 5; ++h) {   // 5 Philox calls -> 10 (x, y) pairs: 8 hazards, vase, spare
            uint32_t q[4];
            Philox::gen(env, ep, 1u + h, 0u, seed, KEY_RESET, q);
            if (h < 4) {
                st[9 + 4 * h] = xm(usym(q[0]), ARENA); st[10 + 4 * h] = xm(usym(q[1]), ARENA);
                st[11 + 4 * h] = xm(usym(q[2]), ARENA); st[12 + 4 * h] = xm(usym(q[3]), ARENA);
            } else {
                st[25] = xm(usym(q[0]), ARENA); st[26] = xm(usym(q[1]), ARENA);
            }
        }
        st[27] = 0.0f;
    }
    __device__ static void observe(const float* st, float* o) {
        using namespace pgoal;
        const float x = st[0], y = st[1], c = st[2], s = st[3], v = st[4], w = st[5];
        // 12 proprioceptive channels
        o[0] = xd(xs(v, st[27]), DT); o[1] = xm(v, w); o[2] = 9.81f;      // accelerometer
        o[3] = v; o[4] = 0.0f; o[5] = 0.0f;                                // velocimeter (body frame)
        o[6] = 0.0f; o[7] = 0.0f; o[8] = w;                                // gyro
        o[9] = c; o[10] = xs(0.0f, s); o[11] = 0.0f;                       // magnetometer
        float* gl = o + 12; float* hl = o + 28; float* vl = o + 44;
#pragma unroll
        for (int k = 0; k < 16; ++k) { gl[k] = 0.0f; hl[k] = 0.0f; vl[k] = 0.0f; }
        // world -> robot frame: rx = c*dx + s*dy ; ry = -s*dx + c*dy
        {
            const float dx = xs(st[6], x), dy = xs(st[7], y);
            lidar_add(gl, xa(xm(c, dx), xm(s, dy)), xs(xm(c, dy), xm(s, dx)));
        }
#pragma unroll
        for (int h = 0; h < NHAZ; ++h) {
            const float dx = xs(st[9 + 2 * h], x), dy = xs(st[10 + 2 * h], y);
            lidar_add(hl, xa(xm(c, dx), xm(s, dy)), xs(xm(c, dy), xm(s, dx)));
        }
        {
            const float dx = xs(st[25], x), dy = xs(st[26], y);
            lidar_add(vl, xa(xm(c, dx), xm(s, dy)), xs(xm(c, dy), xm(s, dx)));
        }
    }
    __device__ static void step(float* st, const float* a, uint32_t seed, uint32_t env, uint32_t ep,
                                float& rew, float& cost, bool& term) {
        using namespace pgoal;
        const float dxo = xs(st[6], st[0]), dyo = xs(st[7], st[1]);
        const float dist_old = xq(xa(xm(dxo, dxo), xm(dyo, dyo)));
        st[27] = st[4];
        car_advance(st, a[0], a[1], VMAX, WMAX, AV, AW, DT);
        // keep the robot inside the arena walls
        st[0] = fminf(ARENA, fmaxf(-ARENA, st[0]));
        st[1] = fminf(ARENA, fmaxf(-ARENA, st[1]));
        const float dxn = xs(st[6], st[0]), dyn = xs(st[7], st[1]);
        const float dist = xq(xa(xm(dxn, dxn), xm(dyn, dyn)));
        rew = xs(dist_old, dist);
        if (dist <= GOAL_R) {
            rew = xa(rew, 1.0f);
            st[8] = xa(st[8], 1.0f);
            sample_goal(st, seed, env, ep, 16u + (uint32_t)st[8]);
        }
        cost = 0.0f;
#pragma unroll
        for (int h = 0; h < NHAZ; ++h) {
            const float dx = xs(st[9 + 2 * h], st[0]), dy = xs(st[10 + 2 * h], st[1]);
            if (xa(xm(dx, dx), xm(dy, dy)) <= HAZ_R * HAZ_R) cost = 1.0f;
        }
        term = false;
    }
};

}  // namespace fsrl
