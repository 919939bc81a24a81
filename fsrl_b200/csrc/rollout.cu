// Rollout collection on the device: one launch == one vector step of
// FastCollector.collect (reference: /root/reference/fsrl/data/fast_collector.py:252-368):
//   policy forward      (:267-269 -> base_policy.py:178-190, tianshou ActorProb/Actor)
//   exploration noise   (:279-280 -> ddpg_lag.py:225-231)
//   map_action          (:284     -> base_policy.py:226-256; remapped action NOT stored)
//   env.step            (:286)    -> envs.cuh (our analytic models)
//   cost / buffer.add   (:325-335) -> SoA transition buffers, env-major, per-env ring
//   done bookkeeping    (:340-363) -> inline when n_episode <= n_env, else rollout_resolve
// fused into a single kernel per step: the actor MLP forward for a tile of envs (mlp.cuh),
// Gaussian sampling from a Philox stream, log-prob, clip/scale, the env step and the SoA
// stores, so no observation or action ever leaves the GPU.
#include "envs.cuh"
#include "mlp.cuh"
#include "fsrl_b200.h"

namespace fsrl {

static_assert(sizeof(fsrl_mlp3_t) == sizeof(Mlp3), "ABI struct mismatch");

__device__ __forceinline__ void gauss_pair(uint32_t a, uint32_t b, float& n0, float& n1) {
    // Box-Muller in f64 (oracle/philox.py normal_pair)
    const double u1 = ((double)a + 1.0) * (1.0 / 4294967296.0);
    const double u2 = (double)b * (1.0 / 4294967296.0);
    const double r = sqrt(-2.0 * log(u1));
    const double ang = 2.0 * 3.141592653589793 * u2;
    n0 = (float)(r * cos(ang));
    n1 = (float)(r * sin(ang));
}

constexpr float LOG_SQRT_2PI = 0.9189385332046727f;

template <int KIND, int H>
__global__ void __launch_bounds__(MLP_TPB)
rollout_step_kernel(const fsrl_rollout_t a) {
    using E_ = Env<KIND>;
    using TT = MlpTile<H>;
    constexpr int D = E_::D, A = E_::A, S = E_::S;
    extern __shared__ __align__(16) float smem[];
    fsrl_collect_stats_t* st = a.stats;
    if (st->finished) return;

    const int tid = threadIdx.x;
    const int e0 = blockIdx.x * TT::R;
    const Mlp3& actor = *reinterpret_cast<const Mlp3*>(&a.actor);
    const MlpSmem<H> sm(smem, D, actor.out);
    constexpr int INP = TT::in_pad(D);
    float* xtile = sm.x;

    // tile-level early out: nothing active in this tile
    __shared__ int s_any;
    if (tid == 0) s_any = 0;
    __syncthreads();
    if (tid < TT::R) {
        const int e = e0 + tid;
        if (e < a.E && a.active[e]) s_any = 1;
    }
    __syncthreads();
    if (!s_any) return;

    // ---- stage the observation tile ---------------------------------------------------------
    mlp_stage_rows<H>(sm, D, [&](int r) -> const float* {
        const int e = e0 + r;
        return e < a.E ? a.obs_cur + (size_t)e * D : nullptr;
    });
    __syncthreads();

    float out[MLP_MAX_OUT];
    if (a.mode != FSRL_MODE_RANDOM) {
        mlp_hidden_forward<H>(actor, sm);
        mlp_head_forward<H>(actor, sm, out);
    }

    // ---- one thread per env: sample, log-prob, map, step, store ------------------------------
    const int r = tid / TT::PARTS, part = tid % TT::PARTS;
    const int e = e0 + r;
    if (part != 0 || e >= a.E || !a.active[e]) return;

    float act[A], mu[A], sig[A], aenv[A];
    float logp = 0.f;
    const uint32_t ctr = a.act_ctr[e];
    float eps[(A + 3) / 4 * 4];
    if (a.mode == FSRL_MODE_TRAIN || a.mode == FSRL_MODE_RANDOM) {
#pragma unroll
        for (int c = 0; c < (A + 3) / 4; ++c) {
            uint32_t rr[4];
            Philox::gen((uint32_t)e, ctr, (uint32_t)c, 0u, a.seed_act, KEY_ACT, rr);
            if (a.mode == FSRL_MODE_RANDOM) {
#pragma unroll
                for (int j = 0; j < 4; ++j) eps[4 * c + j] = usym(rr[j]);   // uniform in [-1, 1)
            } else {
                gauss_pair(rr[0], rr[1], eps[4 * c], eps[4 * c + 1]);
                gauss_pair(rr[2], rr[3], eps[4 * c + 2], eps[4 * c + 3]);
            }
        }
        a.act_ctr[e] = ctr + 1u;
    }
#pragma unroll
    for (int j = 0; j < A; ++j) {
        if (a.mode == FSRL_MODE_RANDOM) {
            // action_space.sample() then map_action_inverse (fast_collector.py:258-264):
            // uniform in [low, high] maps to uniform in [-1, 1] under scaling
            float v = eps[j];
            if (a.action_bound == FSRL_BOUND_TANH) v = 0.5f * (log1pf(v) - log1pf(-v));
            act[j] = v; mu[j] = 0.f; sig[j] = 1.f;
            continue;
        }
        if (a.head == FSRL_HEAD_GAUSS_INDEP) {
            // tianshou ActorProb, state-independent sigma (collect_dataset.py:199-214)
            mu[j] = a.bounded ? a.max_action * tanhf(out[j]) : out[j];
            sig[j] = expf(__ldg(a.log_sigma + j));
        } else if (a.head == FSRL_HEAD_GAUSS_COND) {
            mu[j] = a.bounded ? a.max_action * tanhf(out[j]) : out[j];
            sig[j] = expf(fminf(fmaxf(out[A + j], a.sigma_min), a.sigma_max));
        } else {   // FSRL_HEAD_DETERMINISTIC (tianshou Actor): max_action * tanh(logits)
            mu[j] = a.max_action * tanhf(out[j]);
            sig[j] = 0.f;
        }
        if (a.mode == FSRL_MODE_EVAL || a.head == FSRL_HEAD_DETERMINISTIC) act[j] = mu[j];
        else act[j] = fmaf(sig[j], eps[j], mu[j]);               // dist.sample()  (:189)
    }
    if (a.head == FSRL_HEAD_GAUSS_COND && a.mode != FSRL_MODE_RANDOM) {
        // SAC (sac_lag.py:147-183): squash, log-prob with the tanh correction
        float lp = 0.f;
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const float z = (a.mode == FSRL_MODE_EVAL) ? 0.f : eps[j];
            lp += -0.5f * z * z - logf(sig[j]) - LOG_SQRT_2PI;
            const float sq = tanhf(act[j]);
            lp -= logf(1.0f - sq * sq + a.tanh_eps);
            act[j] = sq;
        }
        logp = lp;
    } else if (a.head == FSRL_HEAD_GAUSS_INDEP && a.mode != FSRL_MODE_RANDOM) {
        // Independent(Normal(mu, sigma), 1).log_prob(act)  (ppo_lag.py:148)
        float lp = 0.f;
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const float z = (act[j] - mu[j]) / sig[j];
            lp += -0.5f * z * z - logf(sig[j]) - LOG_SQRT_2PI;
        }
        logp = lp;
    }
    if (a.head == FSRL_HEAD_DETERMINISTIC && a.mode == FSRL_MODE_TRAIN && a.expl_sigma > 0.f) {
        // DDPG exploration_noise (ddpg_lag.py:225-231): act + N(0, sigma^2)
#pragma unroll
        for (int j = 0; j < A; ++j) act[j] = fmaf(a.expl_sigma, eps[j], act[j]);
    }
    // ---- map_action (base_policy.py:244-256) ---------------------------------------------------
#pragma unroll
    for (int j = 0; j < A; ++j) {
        float v = act[j];
        if (a.action_bound == FSRL_BOUND_CLIP) v = fminf(1.0f, fmaxf(-1.0f, v));
        else if (a.action_bound == FSRL_BOUND_TANH) v = tanhf(v);
        if (a.action_scaling) v = xa(a.act_low[j], xd(xm(xs(a.act_high[j], a.act_low[j]), xa(v, 1.0f)), 2.0f));
        aenv[j] = v;
    }
    // ---- env.step ---------------------------------------------------------------------------------
    float s[S];
#pragma unroll
    for (int i = 0; i < S; ++i) s[i] = a.env_state[(size_t)i * a.E + e];
    float rew, cost;
    bool term;
    const uint32_t ep = a.ep_idx[e] - 1u;
    E_::step(s, aenv, a.seed_env, (uint32_t)e, ep, rew, cost, term);
    const int t_new = a.env_t[e] + 1;
    const bool trunc = (t_new >= a.max_steps) && !term;
    const bool done = term || trunc;
    float on[D];
    E_::observe(s, on);

    // ---- buffer.add (env-major sub-buffer ring; reserved keys of tianshou's buffer) -----------
    if (a.b_obs) {
        const int ptr = a.b_ptr[e];
        const size_t p = (size_t)e * a.cap + ptr;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a.b_obs[p * D + k] = xtile[r * INP + k];
            a.b_obs_next[p * D + k] = on[k];
        }
#pragma unroll
        for (int j = 0; j < A; ++j) a.b_act[p * A + j] = act[j];
        a.b_rew[p] = rew; a.b_cost[p] = cost; a.b_logp[p] = logp;
        a.b_term[p] = term ? 1 : 0; a.b_trunc[p] = trunc ? 1 : 0;
        a.b_ptr[e] = (ptr + 1 == a.cap) ? 0 : ptr + 1;
        const int len = a.b_len[e];
        if (len < a.cap) a.b_len[e] = len + 1;
    }
    // ---- statistics (:326, :338-348) --------------------------------------------------------------
    atomicAdd(&st->step_count, 1ull);
    if (cost != 0.f) atomicAdd(&st->total_cost, (double)cost);
    const double er = a.ep_rew[e] + (double)rew;
    const int el = a.ep_len[e] + 1;
    a.ep_rew[e] = er; a.ep_len[e] = el;
    a.env_t[e] = t_new;

    if (done) {
        if (a.inline_done) {
            // n_episode <= ready envs: every finished env is surplus (:357-363) -> retire it
            atomicAdd(&st->sum_ep_rew, er);
            atomicAdd(&st->sum_ep_len, (unsigned long long)el);
            atomicAdd(term ? &st->term_count : &st->trunc_count, 1);
            a.active[e] = 0;
            a.ep_rew[e] = 0.0; a.ep_len[e] = 0;
            const int c = atomicAdd(&st->episode_count, 1) + 1;
            if (c >= st->n_episode) st->finished_next = 1;
        } else {
            a.done_now[e] = term ? 1 : 2;
        }
    }
#pragma unroll
    for (int i = 0; i < S; ++i) a.env_state[(size_t)i * a.E + e] = s[i];
#pragma unroll
    for (int k = 0; k < D; ++k) a.obs_cur[(size_t)e * D + k] = on[k];
}

// Resolve finished episodes in env order (general path, n_episode > n_env): count, retire the
// first `surplus` finished envs (fast_collector.py:357-363), reset the rest (:351).
template <int KIND>
__global__ void __launch_bounds__(1024) rollout_resolve_kernel(const fsrl_rollout_t a) {
    using E_ = Env<KIND>;
    constexpr int D = E_::D, S = E_::S;
    fsrl_collect_stats_t* st = a.stats;
    __shared__ int s_scan[1024];
    __shared__ int s_base, s_total, s_surplus;
    const int tid = threadIdx.x;
    if (st->finished) return;
    if (st->finished_next) {          // inline path signalled completion during the last step
        if (tid == 0) st->finished = 1;
        return;
    }
    if (a.inline_done) return;
    // pass 1: total number of done envs this step
    int local = 0;
    for (int e = tid; e < a.E; e += 1024) local += (a.active[e] && a.done_now[e]) ? 1 : 0;
    s_scan[tid] = local;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) s_scan[tid] += s_scan[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        s_total = s_scan[0];
        const int epc = st->episode_count + s_total;
        int surplus = st->n_ready - (st->n_episode - epc);
        if (surplus < 0) surplus = 0;
        if (surplus > s_total) surplus = s_total;
        s_surplus = surplus;
        s_base = 0;
    }
    __syncthreads();
    const int total = s_total;
    if (total == 0) return;
    const int surplus = s_surplus;
    // pass 2: ordered walk in chunks of 1024 envs; rank = number of done envs with lower id
    for (int c0 = 0; c0 < a.E; c0 += 1024) {
        const int e = c0 + tid;
        const int flag = (e < a.E && a.active[e] && a.done_now[e]) ? 1 : 0;
        s_scan[tid] = flag;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {      // Hillis-Steele inclusive scan
            int v = (tid >= o) ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int rank = s_base + s_scan[tid] - flag;   // exclusive rank among done envs
        if (flag) {
            const bool term = a.done_now[e] == 1;
            atomicAdd(&st->sum_ep_rew, a.ep_rew[e]);
            atomicAdd(&st->sum_ep_len, (unsigned long long)a.ep_len[e]);
            atomicAdd(term ? &st->term_count : &st->trunc_count, 1);
            a.ep_rew[e] = 0.0; a.ep_len[e] = 0; a.done_now[e] = 0;
            if (rank < surplus) {
                a.active[e] = 0;
            } else {
                float s[S], o[D];
                const uint32_t ep = a.ep_idx[e];
                E_::reset(s, a.seed_env, (uint32_t)e, ep);
                a.ep_idx[e] = ep + 1u;
                a.env_t[e] = 0;
                E_::observe(s, o);
                for (int i = 0; i < S; ++i) a.env_state[(size_t)i * a.E + e] = s[i];
                for (int k = 0; k < D; ++k) a.obs_cur[(size_t)e * D + k] = o[k];
            }
        }
        __syncthreads();
        if (tid == 1023) s_base += s_scan[1023];
        __syncthreads();
    }
    if (tid == 0) {
        st->episode_count += total;
        st->n_ready -= surplus;
        if (st->episode_count >= st->n_episode) st->finished = 1;
    }
}

// reset_env (fast_collector.py:131-152): fresh episode in every env; stats untouched
template <int KIND>
__global__ void env_reset_all_kernel(const fsrl_rollout_t a) {
    using E_ = Env<KIND>;
    constexpr int D = E_::D, S = E_::S;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.E) return;
    float s[S], o[D];
    const uint32_t ep = a.ep_idx[e];
    E_::reset(s, a.seed_env, (uint32_t)e, ep);
    a.ep_idx[e] = ep + 1u;
    a.env_t[e] = 0;
    a.ep_rew[e] = 0.0; a.ep_len[e] = 0; a.done_now[e] = 0;
    E_::observe(s, o);
    for (int i = 0; i < S; ++i) a.env_state[(size_t)i * a.E + e] = s[i];
    for (int k = 0; k < D; ++k) a.obs_cur[(size_t)e * D + k] = o[k];
}

// begin a collect: ready envs = first min(E, n_episode) (:235-236), zero the per-collect stats
__global__ void collect_begin_kernel(const fsrl_rollout_t a, int n_episode) {
    fsrl_collect_stats_t* st = a.stats;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int ready = n_episode < a.E ? n_episode : a.E;
    if (e < a.E) { a.active[e] = e < ready ? 1 : 0; a.done_now[e] = 0; a.ep_rew[e] = 0.0; a.ep_len[e] = 0; }
    if (e == 0) {
        st->step_count = 0; st->total_cost = 0.0; st->sum_ep_rew = 0.0; st->sum_ep_len = 0;
        st->episode_count = 0; st->n_episode = n_episode; st->n_ready = ready;
        st->term_count = 0; st->trunc_count = 0; st->finished = 0; st->finished_next = 0;
    }
}

template <int KIND>
static int launch_step_h(const fsrl_rollout_t& a, cudaStream_t s) {
    const int H = a.actor.H;
#define GO(HH)                                                                                   \
    {                                                                                            \
        using TT = MlpTile<HH>;                                                                  \
        const size_t smem = TT::smem_bytes(Env<KIND>::D);                                        \
        static bool attr_done = false;                                                           \
        if (!attr_done) {                                                                        \
            FSRL_CUDA(cudaFuncSetAttribute(rollout_step_kernel<KIND, HH>,                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_done = true;                                                                    \
        }                                                                                        \
        const int grid = (a.E + TT::R - 1) / TT::R;                                              \
        rollout_step_kernel<KIND, HH><<<grid, MLP_TPB, smem, s>>>(a);                            \
    }
    switch (H) {
        case 64: GO(64) break;
        case 128: GO(128) break;
        case 256: GO(256) break;
        case 512: GO(512) break;
        default: set_error("rollout: hidden width %d unsupported (64/128/256/512)", H); return FSRL_EINVAL;
    }
#undef GO
    FSRL_LAUNCH_CHECK();
    rollout_resolve_kernel<KIND><<<1, 1024, 0, s>>>(a);
    FSRL_LAUNCH_CHECK();
    return FSRL_OK;
}

}  // namespace fsrl

using namespace fsrl;

static int check_rollout(const fsrl_rollout_t* a) {
    FSRL_REQUIRE(a != nullptr, "rollout: null descriptor");
    FSRL_REQUIRE(a->kind >= 0 && a->kind < ENV_KIND_COUNT, "rollout: unknown env kind %d", a->kind);
    FSRL_REQUIRE(a->E > 0, "rollout: E must be positive");
    const EnvDims d = env_dims(a->kind);
    FSRL_REQUIRE(a->actor.in == d.D || a->mode == FSRL_MODE_RANDOM, "rollout: actor input dim %d != obs dim %d", a->actor.in, d.D);
    FSRL_REQUIRE(a->env_state && a->obs_cur && a->env_t && a->ep_idx && a->act_ctr && a->active &&
                 a->ep_rew && a->ep_len && a->done_now && a->stats, "rollout: null state pointer");
    return FSRL_OK;
}

#define DISPATCH_KIND(kind, CALL)                                                    \
    switch (kind) {                                                                  \
        case ENV_CAR_CIRCLE: { constexpr int K = ENV_CAR_CIRCLE; CALL; } break;      \
        case ENV_CAR_RUN: { constexpr int K = ENV_CAR_RUN; CALL; } break;            \
        case ENV_BALL_CIRCLE: { constexpr int K = ENV_BALL_CIRCLE; CALL; } break;    \
        case ENV_BALL_RUN: { constexpr int K = ENV_BALL_RUN; CALL; } break;          \
        case ENV_ANT_CIRCLE: { constexpr int K = ENV_ANT_CIRCLE; CALL; } break;      \
        case ENV_POINT_GOAL: { constexpr int K = ENV_POINT_GOAL; CALL; } break;      \
        default: set_error("unknown env kind %d", kind); return FSRL_EINVAL;         \
    }

extern "C" int fsrl_env_dims(int kind, int* D, int* A, int* S, int* T) {
    FSRL_REQUIRE(kind >= 0 && kind < ENV_KIND_COUNT, "fsrl_env_dims: unknown env kind %d", kind);
    const EnvDims d = env_dims(kind);
    if (D) *D = d.D; if (A) *A = d.A; if (S) *S = d.S; if (T) *T = d.T;
    return FSRL_OK;
}

extern "C" int fsrl_env_reset_all(const fsrl_rollout_t* a, void* stream) {
    int rc = check_rollout(a);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH_KIND(a->kind, (env_reset_all_kernel<K><<<(a->E + 127) / 128, 128, 0, s>>>(*a)));
    FSRL_LAUNCH_CHECK();
    return FSRL_OK;
}

extern "C" int fsrl_collect_begin(const fsrl_rollout_t* a, int n_episode, void* stream) {
    int rc = check_rollout(a);
    if (rc) return rc;
    FSRL_REQUIRE(n_episode > 0, "n_episode must be positive");   // fast_collector.py:234
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    collect_begin_kernel<<<(a->E + 127) / 128, 128, 0, s>>>(*a, n_episode);
    FSRL_LAUNCH_CHECK();
    return FSRL_OK;
}

extern "C" int fsrl_rollout_steps(const fsrl_rollout_t* a, int n_steps, void* stream) {
    int rc = check_rollout(a);
    if (rc) return rc;
    FSRL_REQUIRE(n_steps >= 0, "fsrl_rollout_steps: n_steps < 0");
    FSRL_REQUIRE(a->mode == FSRL_MODE_RANDOM || (a->actor.w1t && a->actor.w2t && a->actor.w3t),
                 "rollout: null actor weights");
    FSRL_REQUIRE(a->actor.out <= MLP_MAX_OUT, "rollout: actor out dim %d > %d", a->actor.out, MLP_MAX_OUT);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < n_steps; ++i) {
        DISPATCH_KIND(a->kind, { int r2 = launch_step_h<K>(*a, s); if (r2) return r2; });
    }
    return FSRL_OK;
}
