// PPO-Lagrangian minibatch update on the device: clipped surrogate + lambda * cost-advantage
// actor loss, value losses for the reward and cost critics, backward pass, global-norm clip
// and Adam -- three launches per minibatch, no host round trip inside a repeat.
//
// Replaces (reference, eager PyTorch + ~10 .item() syncs per minibatch):
//   /root/reference/fsrl/policy/ppo_lag.py:173-212  policy_loss (per-minibatch adv norm
//        :178-182, clipped surrogate :185-193, unclipped cost term :196-198, rescaling
//        :200-201, approx_kl :204)
//   /root/reference/fsrl/policy/ppo_lag.py:152-171  critics_loss
//   /root/reference/fsrl/policy/ppo_lag.py:223-247  forward/backward/clip_grad_norm_/Adam
//   /root/reference/fsrl/policy/lagrangian_base.py:145-166  safety_loss
//
// Phase A (ppo_fwdbwd): grid (row tiles, nets).  Gathers the minibatch rows by permuted
//   index, runs the fused MLP forward (mlp.cuh), evaluates the loss gradient at the head and
//   back-propagates to dZ2 / dZ1; activations needed for the weight gradients go to an
//   L2-resident scratch.
// Phase B (ppo_wgrad): weight gradients as outer-product accumulations over the minibatch,
//   each CTA owning a 32x64 tile of dW2t (no cross-CTA reduction), plus three small CTAs per
//   net for layer 1 / layer 3 / biases; sum of squares for the global norm via one atomic per
//   CTA.
// Phase C (adam): clip scale + Adam over the flat parameter buffer; the W2 blocks are
//   processed in 32x32 tiles through shared memory so that both the canonical W2t and its
//   out-major mirror (needed by the backward GEMM) are written coalesced.
#include "mlp.cuh"
#include "fsrl_b200.h"

namespace fsrl {

constexpr int ST_ACTOR_REW = 0, ST_ACTOR_SAFETY = 1, ST_KL = 2, ST_VF0 = 3, ST_VF1 = 4,
              ST_ENTROPY = 5, ST_GRADNORM = 6, ST_CLIPFRAC = 7;
constexpr float LOG_SQRT_2PI_P = 0.9189385332046727f;
constexpr int DOUT_LD = 16;   // scratch row stride of dOut (cols [A, 2A) carry dlog_sigma)

__device__ long long g_dbg_clock[16];
#define DBG_T(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_dbg_clock[i] = clock64(); } while (0)

__device__ __forceinline__ int slot_mb(const fsrl_ppo_update_t& u, int mb_off) { return mb_off / u.batch_size; }

struct NetView {   // resolved pointers of one network inside the flat buffers
    Mlp3 m;
    const float* w2n;      // mirror [out][in] of w2t
    const float* log_sigma;
    float *g_w1t, *g_b1, *g_w2t, *g_b2, *g_w3t, *g_b3, *g_log_sigma;
    float *s_h1, *s_h2, *s_dz1, *s_dz2, *s_dout;   // scratch [Bmax][H] / [Bmax][16]
};

__device__ __forceinline__ NetView net_view(const fsrl_ppo_update_t& u, int n) {
    NetView v;
    const int H = u.H, D = u.D;
    const int out = (n == 0) ? u.actor_out : 1;
    const float* th = u.theta + u.net_off[n];
    float* g = u.grad + u.net_off[n];
    size_t o = 0;
    v.m.w1t = th + o; v.g_w1t = g + o; o += (size_t)D * H;
    v.m.b1 = th + o;  v.g_b1 = g + o;  o += H;
    v.m.w2t = th + o; v.g_w2t = g + o; o += (size_t)H * H;
    v.m.b2 = th + o;  v.g_b2 = g + o;  o += H;
    v.m.w3t = th + o; v.g_w3t = g + o; o += (size_t)H * out;
    v.m.b3 = th + o;  v.g_b3 = g + o;  o += out;
    v.log_sigma = th + o; v.g_log_sigma = g + o;
    v.m.in = D; v.m.H = H; v.m.out = out;
    v.w2n = u.w2n + (size_t)n * H * H;
    float* sc = u.scratch + (size_t)n * u.bmax * (4 * (size_t)H + DOUT_LD);
    v.s_h1 = sc; v.s_h2 = sc + (size_t)u.bmax * H; v.s_dz1 = sc + 2 * (size_t)u.bmax * H;
    v.s_dz2 = sc + 3 * (size_t)u.bmax * H; v.s_dout = sc + 4 * (size_t)u.bmax * H;
    return v;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < MLP_TPB / 32; ++i) t += red[i];
    return t;
}

// ------------------------------------------------------------------------------------------
// Phase A
// ------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(MLP_TPB)
ppo_fwdbwd_kernel(const fsrl_ppo_update_t u, int mb_off, int B, int slot) {
    using TT = MlpTile<H>;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int net = blockIdx.y;
    const int r0 = blockIdx.x * TT::R;
    const int D = u.D;
    const NetView nv = net_view(u, net);
    const MlpSmem<H> sm(smem, D, nv.m.out);
    float* dz = sm.end(nv.m.out);                     // [R][LDA]
    float* sdout = dz + (size_t)TT::R * TT::LDA;      // [R][DOUT_LD]
    __shared__ int s_idx[64];
    __shared__ float s_red[MLP_TPB / 32];
    __shared__ float s_mean[2], s_rstd[2];

    DBG_T(0);
    const int* perm = u.perm + mb_off;
    if (net == 0 && blockIdx.x == 0 && tid == 0) *u.norm_sq = 0.f;   // consumed by phase C of the previous step

    if (tid < TT::R) s_idx[tid] = (r0 + tid < B) ? perm[r0 + tid] : -1;
    __syncthreads();
    mlp_stage_rows<H>(sm, D, [&](int r) -> const float* {
        const int id = s_idx[r];
        return id >= 0 ? u.obs + (size_t)id * D : nullptr;
    });
    // per-minibatch advantage normalisation (ppo_lag.py:178-182): mean, unbiased std, no eps
    if (net == 0 && u.moments != nullptr) {
        // data-parallel run: sum / sum of squares of this minibatch were reduced over all ranks
        // beforehand (ppo_adv_moments_kernel + one all-reduce per repeat)
        if (tid < u.C) {
            const double* mo = u.moments + ((size_t)slot_mb(u, mb_off) * 2 + tid) * 2;
            const double nn = (double)B * (double)u.world;
            const double mean = mo[0] / nn;
            const double var = (mo[1] - nn * mean * mean) / (nn - 1.0);
            s_mean[tid] = u.norm_adv ? (float)mean : 0.f;
            s_rstd[tid] = u.norm_adv ? (float)(1.0 / sqrt(var)) : 1.0f;
        }
    } else if (net == 0) {
        for (int c = 0; c < u.C; ++c) {
            float s = 0.f;
            for (int i = tid; i < B; i += MLP_TPB) s += u.adv[(size_t)c * u.ld + perm[i]];
            const float mean = block_sum_256(s, s_red) / (float)B;
            float q = 0.f;
            for (int i = tid; i < B; i += MLP_TPB) {
                const float d = u.adv[(size_t)c * u.ld + perm[i]] - mean;
                q += d * d;
            }
            const float var = block_sum_256(q, s_red) / (float)(B - 1);
            if (tid == 0) {
                s_mean[c] = u.norm_adv ? mean : 0.f;
                s_rstd[c] = u.norm_adv ? 1.0f / sqrtf(var) : 1.0f;
            }
        }
    }
    __syncthreads();
    DBG_T(1);

    mlp_hidden_forward<H>(nv.m, sm);
    DBG_T(2);
    float out[MLP_MAX_OUT];
    mlp_head_forward<H>(nv.m, sm, out);
    DBG_T(3);

    // ---- loss gradient at the head: one thread per row -----------------------------------------
    const int r = tid / TT::PARTS, part = tid % TT::PARTS;
    float st_a = 0.f, st_b = 0.f, st_c = 0.f, st_d = 0.f;     // per-thread stat partials
    if (part == 0) {
        float dd[DOUT_LD];
#pragma unroll
        for (int j = 0; j < DOUT_LD; ++j) dd[j] = 0.f;
        const int id = s_idx[r];
        if (id >= 0) {
            const float invB = 1.0f / (float)B;
            if (net == 0) {
                const int A = u.A;
                float logp = 0.f, zz[8], sg[8], dmu[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < A) {
                        const float t = tanhf(out[j]);
                        const float mu = u.bounded ? u.max_action * t : out[j];
                        dmu[j] = u.bounded ? u.max_action * (1.0f - t * t) : 1.0f;
                        sg[j] = expf(nv.log_sigma[j]);
                        zz[j] = (u.act[(size_t)id * A + j] - mu) / sg[j];
                        logp += -0.5f * zz[j] * zz[j] - nv.log_sigma[j] - LOG_SQRT_2PI_P;
                    }
                }
                const float lpo = u.logp_old[id];
                const float ratio = expf(logp - lpo);
                const float ar = (u.adv[id] - s_mean[0]) * s_rstd[0];
                const float surr1 = ratio * ar;
                const float rc = fminf(fmaxf(ratio, 1.0f - u.eps_clip), 1.0f + u.eps_clip);
                const float surr2 = rc * ar;
                // d(-min(surr1, surr2))/d ratio ; ties split evenly like torch.min's backward
                const bool inside = (ratio >= 1.0f - u.eps_clip) && (ratio <= 1.0f + u.eps_clip);
                float g_ratio;   // d loss_rew_i / d ratio  (before the 1/B of the mean)
                float lrew;
                if (surr1 < surr2) { g_ratio = -ar; lrew = -surr1; }
                else if (surr1 > surr2) { g_ratio = inside ? -ar : 0.f; lrew = -surr2; }
                else { g_ratio = inside ? -ar : -0.5f * ar; lrew = -surr1; }
                if (u.dual_clip > 0.f && ar < 0.f) {
                    // clip2 = max(min(s1,s2), dual_clip*adv) for negative advantages (:188-191)
                    const float c1 = fminf(surr1, surr2), c2 = u.dual_clip * ar;
                    if (c2 > c1) { g_ratio = 0.f; lrew = -c2; }
                    else if (c2 == c1) { g_ratio *= 0.5f; }
                }
                float g_saf = 0.f, lsaf = 0.f;
                if (u.use_lagrangian && u.C > 1) {
                    const float ac = (u.adv[(size_t)u.ld + id] - s_mean[1]) * s_rstd[1];
                    g_saf = ac * u.lagrangian;          // d mean(ratio*adv_c*lambda) / d ratio
                    lsaf = ratio * ac * u.lagrangian;
                }
                // d loss / d logp = rescaling * (g_ratio + g_saf) * ratio / B
                const float gl = u.rescaling * (g_ratio + g_saf) * ratio * invB;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < A) {
                        dd[j] = gl * (zz[j] / sg[j]) * dmu[j];        // via mu
                        dd[A + j] = gl * (zz[j] * zz[j] - 1.0f);       // via log_sigma
                    }
                }
                st_a = lrew * invB; st_b = lsaf * invB; st_c = (lpo - logp) * invB;
            } else {
                const int c = net - 1;
                const float v = out[0];
                const float ret = u.ret[(size_t)c * u.ld + id];
                float lv, gv;
                if (u.value_clip) {
                    const float vt = u.values[(size_t)c * u.ld + id];
                    const float dv = fminf(fmaxf(v - vt, -u.eps_clip), u.eps_clip);
                    const float vc = vt + dv;
                    const float vf1 = (ret - v) * (ret - v), vf2 = (ret - vc) * (ret - vc);
                    const bool in_clip = (v - vt > -u.eps_clip) && (v - vt < u.eps_clip);
                    if (vf1 > vf2) { lv = vf1; gv = 2.0f * (v - ret); }
                    else if (vf1 < vf2) { lv = vf2; gv = in_clip ? 2.0f * (vc - ret) : 0.f; }
                    else { lv = vf1; gv = in_clip ? 2.0f * (v - ret) : (v - ret); }
                } else {
                    lv = (ret - v) * (ret - v);
                    gv = 2.0f * (v - ret);
                }
                dd[0] = u.vf_coef * gv * invB;
                st_d = lv * invB;
            }
        }
#pragma unroll
        for (int j = 0; j < DOUT_LD; ++j) sdout[r * DOUT_LD + j] = dd[j];
        if (r0 + r < u.bmax) {
#pragma unroll
            for (int j = 0; j < DOUT_LD; j += 4)
                *reinterpret_cast<float4*>(nv.s_dout + (size_t)(r0 + r) * DOUT_LD + j) =
                    make_float4(dd[j], dd[j + 1], dd[j + 2], dd[j + 3]);
        }
    }
    DBG_T(4);
    // minibatch statistics (loss/actor_rew, actor_safety, kl, vf_i): one atomic per CTA each
    {
        float* stat = u.stats + (size_t)slot * FSRL_PPO_STATS;
        if (net == 0) {
            const float a = block_sum_256(st_a, s_red), b = block_sum_256(st_b, s_red), c = block_sum_256(st_c, s_red);
            if (tid == 0) {
                atomicAdd(stat + ST_ACTOR_REW, a); atomicAdd(stat + ST_ACTOR_SAFETY, b); atomicAdd(stat + ST_KL, c);
                if (blockIdx.x == 0) {
                    float ent = 0.f;
                    for (int j = 0; j < u.A; ++j) ent += 0.5f + LOG_SQRT_2PI_P + nv.log_sigma[j];
                    stat[ST_ENTROPY] = ent;
                }
            }
        } else {
            const float d = block_sum_256(st_d, s_red);
            if (tid == 0) atomicAdd(stat + ST_VF0 + (net - 1), d);
        }
    }
    __syncthreads();

    DBG_T(5);
    // ---- backward through layer 3 and ReLU 2; spill h1 / h2 / dz2 for the weight gradients ------
    const int nout = (net == 0) ? u.A : 1;       // head columns that feed w3t (mu only)
    const int wout = nv.m.out;
    for (int e = tid; e < TT::R * (H / 4); e += MLP_TPB) {
        const int row = e / (H / 4), k4 = (e % (H / 4)) * 4;
        float a4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < nout; ++j) {
            const float g = sdout[row * DOUT_LD + j];
#pragma unroll
            for (int q = 0; q < 4; ++q) a4[q] = fmaf(g, sm.w3s[(size_t)(k4 + q) * wout + j], a4[q]);
        }
        const float4 hv = *reinterpret_cast<const float4*>(sm.h2 + (size_t)row * TT::LDA + k4);
        const float4 g4 = make_float4(hv.x > 0.f ? a4[0] : 0.f, hv.y > 0.f ? a4[1] : 0.f,
                                      hv.z > 0.f ? a4[2] : 0.f, hv.w > 0.f ? a4[3] : 0.f);
        *reinterpret_cast<float4*>(dz + (size_t)row * TT::LDA + k4) = g4;
        if (r0 + row < u.bmax) {
            *reinterpret_cast<float4*>(nv.s_dz2 + (size_t)(r0 + row) * H + k4) = g4;
            *reinterpret_cast<float4*>(nv.s_h2 + (size_t)(r0 + row) * H + k4) = hv;
            *reinterpret_cast<float4*>(nv.s_h1 + (size_t)(r0 + row) * H + k4) =
                *reinterpret_cast<const float4*>(sm.h1 + (size_t)row * TT::LDA + k4);
        }
    }
    DBG_T(6);
    // ---- backward through layer 2: dH1 = dZ2 . W2 (W2n is [out][in]) then ReLU 1 ------------------
    {
        float c[TT::MT][TT::NT][4];
        tc_init_bias<H>(c, nullptr);
        tc_gemm<H>(c, dz, TT::LDA, H, nv.w2n, sm.wst, false);
        tc_foreach<H>(c, [&](int row, int col, float v0, float v1) {
            if (r0 + row < u.bmax) {
                const float2 hv = *reinterpret_cast<const float2*>(sm.h1 + (size_t)row * TT::LDA + col);
                *reinterpret_cast<float2*>(nv.s_dz1 + (size_t)(r0 + row) * H + col) =
                    make_float2(hv.x > 0.f ? v0 : 0.f, hv.y > 0.f ? v1 : 0.f);
            }
        });
    }
    DBG_T(7);
}

// ------------------------------------------------------------------------------------------
// Phase B: weight gradients
// ------------------------------------------------------------------------------------------
constexpr int WG_TPB = 128, WG_TK = 32, WG_TO = 64, WG_RC = 32;

__device__ __forceinline__ float block_sum_128(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}

// Roles by blockIdx.x (per net):  [0, NT) dW2t tiles (32 k x 64 o; k-tile 0 also emits db2)
//                                 [NT, NT+NTO) layer 1: dW1t[:, o-tile], db1[o-tile]
//                                 [NT+NTO, NT+2*NTO) layer 3: dW3t[k-tile, :] (+ db3, dlog_sigma)
// Every role streams the minibatch in chunks of 32 rows: the next chunk is prefetched into
// registers while the current one is consumed from shared memory.
template <int H>
__global__ void __launch_bounds__(WG_TPB)
ppo_wgrad_kernel(const fsrl_ppo_update_t u, int mb_off, int B) {
    constexpr int NTK = H / WG_TK, NTO = H / WG_TO, NT = NTK * NTO;
    __shared__ __align__(16) float sL[WG_RC][WG_TO];     // left operand chunk (<= 64 cols)
    __shared__ __align__(16) float sG[WG_RC][WG_TO];     // right operand chunk
    __shared__ float s_red[4];
    const int tid = threadIdx.x;
    const int net = blockIdx.y;
    const NetView nv = net_view(u, net);
    const int bx = blockIdx.x;
    const int nchunk = (B + WG_RC - 1) / WG_RC;
    float sq = 0.f;
    if (bx < NT) {
        // ---- dW2t[k][o] = sum_r h1[r][k] * dz2[r][o] : 32 x 64 tile, 4 x 4 per thread ----------
        const int k0 = (bx / NTO) * WG_TK, o0 = (bx % NTO) * WG_TO;
        const int tk = tid / 16, to = tid % 16;
        const bool do_bias = (k0 == 0);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
        float bsum = 0.f;                                  // db2 column sum (threads < 64)
        float4 pl[2], pg[4];
        auto prefetch = [&](int rb) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int f = tid + q * WG_TPB, rr = f / 8, cc = (f % 8) * 4;
                pl[q] = (rb + rr < B) ? __ldcg(reinterpret_cast<const float4*>(nv.s_h1 + (size_t)(rb + rr) * H + k0 + cc))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int f = tid + q * WG_TPB, rr = f / 16, cc = (f % 16) * 4;
                pg[q] = (rb + rr < B) ? __ldcg(reinterpret_cast<const float4*>(nv.s_dz2 + (size_t)(rb + rr) * H + o0 + cc))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        prefetch(0);
        for (int ch = 0; ch < nchunk; ++ch) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int f = tid + q * WG_TPB, rr = f / 8, cc = (f % 8) * 4;
                *reinterpret_cast<float4*>(&sL[rr][cc]) = pl[q];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int f = tid + q * WG_TPB, rr = f / 16, cc = (f % 16) * 4;
                *reinterpret_cast<float4*>(&sG[rr][cc]) = pg[q];
            }
            __syncthreads();
            if (ch + 1 < nchunk) prefetch((ch + 1) * WG_RC);
#pragma unroll 8
            for (int rr = 0; rr < WG_RC; ++rr) {
                const float4 l = *reinterpret_cast<const float4*>(&sL[rr][4 * tk]);
                const float4 g = *reinterpret_cast<const float4*>(&sG[rr][4 * to]);
                acc[0][0] = fmaf(l.x, g.x, acc[0][0]); acc[0][1] = fmaf(l.x, g.y, acc[0][1]);
                acc[0][2] = fmaf(l.x, g.z, acc[0][2]); acc[0][3] = fmaf(l.x, g.w, acc[0][3]);
                acc[1][0] = fmaf(l.y, g.x, acc[1][0]); acc[1][1] = fmaf(l.y, g.y, acc[1][1]);
                acc[1][2] = fmaf(l.y, g.z, acc[1][2]); acc[1][3] = fmaf(l.y, g.w, acc[1][3]);
                acc[2][0] = fmaf(l.z, g.x, acc[2][0]); acc[2][1] = fmaf(l.z, g.y, acc[2][1]);
                acc[2][2] = fmaf(l.z, g.z, acc[2][2]); acc[2][3] = fmaf(l.z, g.w, acc[2][3]);
                acc[3][0] = fmaf(l.w, g.x, acc[3][0]); acc[3][1] = fmaf(l.w, g.y, acc[3][1]);
                acc[3][2] = fmaf(l.w, g.z, acc[3][2]); acc[3][3] = fmaf(l.w, g.w, acc[3][3]);
            }
            if (do_bias && tid < WG_TO) {
#pragma unroll 8
                for (int rr = 0; rr < WG_RC; ++rr) bsum += sG[rr][tid];
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<float4*>(nv.g_w2t + (size_t)(k0 + 4 * tk + i) * H + o0 + 4 * to) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            sq += acc[i][0] * acc[i][0] + acc[i][1] * acc[i][1] + acc[i][2] * acc[i][2] + acc[i][3] * acc[i][3];
        }
        if (do_bias && tid < WG_TO) { nv.g_b2[o0 + tid] = bsum; sq += bsum * bsum; }
    } else if (bx < NT + NTO) {
        // ---- layer 1: dW1t[d][o] = sum_r x[r][d] * dz1[r][o];  db1[o] = sum_r dz1[r][o] ---------
        const int D = u.D;
        const int o0 = (bx - NT) * WG_TO;
        const int* perm = u.perm + mb_off;
        const int o = tid % WG_TO, dg = tid / WG_TO;        // 2 d-groups of 8 per pass
        for (int d0 = 0; d0 < D; d0 += 16) {
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.f;
            float bsum = 0.f;
            float4 pg[4];
            float px[4];
            auto prefetch = [&](int rb) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int f = tid + q * WG_TPB, rr = f / 16, cc = (f % 16) * 4;
                    pg[q] = (rb + rr < B) ? __ldcg(reinterpret_cast<const float4*>(nv.s_dz1 + (size_t)(rb + rr) * H + o0 + cc))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {       // x chunk: 32 rows x 16 d  = 512 values
                    const int f = tid + q * WG_TPB, rr = f / 16, dd = d0 + (f % 16);
                    px[q] = (rb + rr < B && dd < D) ? __ldg(u.obs + (size_t)perm[rb + rr] * D + dd) : 0.f;
                }
            };
            prefetch(0);
            for (int ch = 0; ch < nchunk; ++ch) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int f = tid + q * WG_TPB;
                    *reinterpret_cast<float4*>(&sG[f / 16][(f % 16) * 4]) = pg[q];
                    sL[f / 16][f % 16] = px[q];
                }
                __syncthreads();
                if (ch + 1 < nchunk) prefetch((ch + 1) * WG_RC);
#pragma unroll 8
                for (int rr = 0; rr < WG_RC; ++rr) {
                    const float g = sG[rr][o];
                    const float4 xa4 = *reinterpret_cast<const float4*>(&sL[rr][8 * dg]);
                    const float4 xb4 = *reinterpret_cast<const float4*>(&sL[rr][8 * dg + 4]);
                    acc[0] = fmaf(xa4.x, g, acc[0]); acc[1] = fmaf(xa4.y, g, acc[1]);
                    acc[2] = fmaf(xa4.z, g, acc[2]); acc[3] = fmaf(xa4.w, g, acc[3]);
                    acc[4] = fmaf(xb4.x, g, acc[4]); acc[5] = fmaf(xb4.y, g, acc[5]);
                    acc[6] = fmaf(xb4.z, g, acc[6]); acc[7] = fmaf(xb4.w, g, acc[7]);
                    bsum += g;
                }
                __syncthreads();
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int d = d0 + 8 * dg + q;
                if (d < D) { nv.g_w1t[(size_t)d * H + o0 + o] = acc[q]; sq += acc[q] * acc[q]; }
            }
            if (d0 == 0 && dg == 0) { nv.g_b1[o0 + o] = bsum; sq += bsum * bsum; }
        }
    } else {
        // ---- layer 3: dW3t[k][j] = sum_r h2[r][k] * dout[r][j];  db3;  dlog_sigma ------------------
        const int out = nv.m.out;
        const int A = u.A;
        const int k0 = (bx - NT - NTO) * WG_TO;
        const int k = tid % WG_TO, jg = tid / WG_TO;         // j = 8*jg + q
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        float csum = 0.f;                                     // column sum of dout (threads < 16)
        float4 pg[4];
        float4 pd;
        auto prefetch = [&](int rb) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int f = tid + q * WG_TPB, rr = f / 16, cc = (f % 16) * 4;
                pg[q] = (rb + rr < B) ? __ldcg(reinterpret_cast<const float4*>(nv.s_h2 + (size_t)(rb + rr) * H + k0 + cc))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const int rr = tid / 4, cc = (tid % 4) * 4;      // dout chunk: 32 rows x 16
            pd = (rb + rr < B) ? __ldcg(reinterpret_cast<const float4*>(nv.s_dout + (size_t)(rb + rr) * DOUT_LD + cc))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        prefetch(0);
        for (int ch = 0; ch < nchunk; ++ch) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int f = tid + q * WG_TPB;
                *reinterpret_cast<float4*>(&sG[f / 16][(f % 16) * 4]) = pg[q];
            }
            *reinterpret_cast<float4*>(&sL[tid / 4][(tid % 4) * 4]) = pd;
            __syncthreads();
            if (ch + 1 < nchunk) prefetch((ch + 1) * WG_RC);
#pragma unroll 8
            for (int rr = 0; rr < WG_RC; ++rr) {
                const float h = sG[rr][k];
                const float4 da = *reinterpret_cast<const float4*>(&sL[rr][8 * jg]);
                const float4 db = *reinterpret_cast<const float4*>(&sL[rr][8 * jg + 4]);
                acc[0] = fmaf(h, da.x, acc[0]); acc[1] = fmaf(h, da.y, acc[1]);
                acc[2] = fmaf(h, da.z, acc[2]); acc[3] = fmaf(h, da.w, acc[3]);
                acc[4] = fmaf(h, db.x, acc[4]); acc[5] = fmaf(h, db.y, acc[5]);
                acc[6] = fmaf(h, db.z, acc[6]); acc[7] = fmaf(h, db.w, acc[7]);
            }
            if (k0 == 0 && tid < DOUT_LD) {
#pragma unroll 8
                for (int rr = 0; rr < WG_RC; ++rr) csum += sL[rr][tid];
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = 8 * jg + q;
            if (j < out) { nv.g_w3t[(size_t)(k0 + k) * out + j] = acc[q]; sq += acc[q] * acc[q]; }
        }
        if (k0 == 0 && tid < DOUT_LD) {
            if (tid < out) { nv.g_b3[tid] = csum; sq += csum * csum; }
            else if (net == 0 && u.head_indep && tid >= A && tid < 2 * A) { nv.g_log_sigma[tid - A] = csum; sq += csum * csum; }
        }
    }
    const float tot = block_sum_128(sq, s_red);
    if (tid == 0 && tot != 0.f) atomicAdd(u.norm_sq, tot);
}

// ------------------------------------------------------------------------------------------
// Phase C: clip_grad_norm_ + Adam (torch.optim.Adam single-tensor arithmetic order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float w1, float b2,
                                          float w2, float bc2s, float eps, float neg_step) {
    m = m + w1 * (g - m);                 // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + (w2 * g) * g;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / bc2s + eps;
    return p + (neg_step * m) / denom;    // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256)
adam_kernel(const fsrl_ppo_update_t u, float w1, float b2, float w2, float bc2s, float eps,
            float neg_step, int slot, int n_plain_blocks) {
    __shared__ float tile[32][33];
    const float gs = (u.world > 1) ? 1.0f / (float)u.world : 1.0f;     // average the summed gradients
    float scale = gs;
    const float nsq = *u.norm_sq * gs * gs;
    if (u.max_grad_norm > 0.f) {
        const float coef = u.max_grad_norm / (sqrtf(nsq) + 1e-6f);
        scale = gs * fminf(coef, 1.0f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && u.stats && slot >= 0)
        u.stats[(size_t)slot * FSRL_PPO_STATS + ST_GRADNORM] = sqrtf(nsq);
    const int H = u.H;
    if ((int)blockIdx.x < n_plain_blocks) {
        // everything except the W2 blocks (they are skipped here by range test)
        const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
        if (i >= u.n_params) return;
        for (int n = 0; n < u.n_nets; ++n) {
            const long long w2s = u.net_off[n] + (long long)u.D * H + H;
            if (i >= w2s && i < w2s + (long long)H * H) return;
        }
        float m = u.adam_m[i], v = u.adam_v[i];
        const float g = (u.mask && u.mask[i] == 0) ? 0.f : u.grad[i] * scale;
        if (u.mask && u.mask[i] == 0) return;
        u.theta[i] = adam_one(u.theta[i], g, m, v, w1, b2, w2, bc2s, eps, neg_step);
        u.adam_m[i] = m; u.adam_v[i] = v;
    } else {
        // W2 tiles: 32 x 32, update canonical W2t[k][o] and its mirror W2n[o][k]
        const int tpn = (H / 32) * (H / 32);
        const int t = blockIdx.x - n_plain_blocks;
        const int n = t / tpn, tt = t % tpn;
        const int k0 = (tt / (H / 32)) * 32, o0 = (tt % (H / 32)) * 32;
        const long long base = u.net_off[n] + (long long)u.D * H + H;
        const int lx = threadIdx.x % 32, ly = threadIdx.x / 32;
        const bool frozen = u.mask && u.mask[base] == 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int kk = ly + 8 * q;
            const long long i = base + (long long)(k0 + kk) * H + o0 + lx;
            float p = u.theta[i];
            if (!frozen) {
                float m = u.adam_m[i], v = u.adam_v[i];
                p = adam_one(p, u.grad[i] * scale, m, v, w1, b2, w2, bc2s, eps, neg_step);
                u.theta[i] = p; u.adam_m[i] = m; u.adam_v[i] = v;
            }
            tile[kk][lx] = p;
        }
        __syncthreads();
        float* mir = u.w2n + (size_t)n * H * H;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int oo = ly + 8 * q;
            mir[(size_t)(o0 + oo) * H + k0 + lx] = tile[lx][oo];
        }
    }
}

// sum of squares of the (all-reduced) gradient buffer -> *u.norm_sq
__global__ void __launch_bounds__(1024) grad_norm_kernel(const fsrl_ppo_update_t u) {
    __shared__ float red[32];
    float s = 0.f;
    for (long long i = threadIdx.x; i < u.n_params; i += 1024) { const float g = u.grad[i]; s += g * g; }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 32; ++w) t += red[w]; *u.norm_sq = t; }
}

// per-minibatch sum and sum of squares of the advantages (block b = minibatch b of the repeat)
__global__ void __launch_bounds__(256) ppo_adv_moments_kernel(const fsrl_ppo_update_t u, long long n_total, int n_mb) {
    __shared__ double red[2][8];
    const int mb = blockIdx.x;
    const long long off = (long long)mb * u.batch_size;
    long long B = u.batch_size;
    if (mb == n_mb - 1) B = n_total - off;
    for (int c = 0; c < u.C; ++c) {
        double s = 0.0, q = 0.0;
        for (long long i = threadIdx.x; i < B; i += 256) {
            const double a = (double)u.adv[(size_t)c * u.ld + u.perm[off + i]];
            s += a; q += a * a;
        }
        s = warp_sum(s); q = warp_sum(q);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = q; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ts = 0.0, tq = 0.0;
            for (int w = 0; w < 8; ++w) { ts += red[0][w]; tq += red[1][w]; }
            u.moments_w[((size_t)mb * 2 + c) * 2] = ts;
            u.moments_w[((size_t)mb * 2 + c) * 2 + 1] = tq;
        }
    }
}

extern "C" int fsrl_allreduce_fused(void* comm, float* buf, long long n, void* stream);
extern "C" int fsrl_allreduce_f64(void* comm, double* buf, long long n, void* stream);

template <int H>
static int ppo_launch_minibatch(const fsrl_ppo_update_t& u, int mb_off, int B, int slot,
                                long long adam_t, cudaStream_t s) {
    using TT = MlpTile<H>;
    const size_t smemA = TT::smem_bytes(u.D) + sizeof(float) * ((size_t)TT::R * TT::LDA + (size_t)TT::R * DOUT_LD);
    static bool attr_done = false;
    if (!attr_done) {
        FSRL_CUDA(cudaFuncSetAttribute(ppo_fwdbwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
        attr_done = true;
    }
    const dim3 gA((B + TT::R - 1) / TT::R, u.n_nets);
    ppo_fwdbwd_kernel<H><<<gA, MLP_TPB, smemA, s>>>(u, mb_off, B, slot);
    FSRL_LAUNCH_CHECK();
    const dim3 gB((H / WG_TK) * (H / WG_TO) + 2 * (H / WG_TO), u.n_nets);
    ppo_wgrad_kernel<H><<<gB, WG_TPB, 0, s>>>(u, mb_off, B);
    FSRL_LAUNCH_CHECK();
    if (u.world > 1) {
        // data parallel: ONE all-reduce of the flat gradient buffer per optimiser step, then the
        // global norm of the reduced gradient (the local partial norms are meaningless now)
        int rc = fsrl_allreduce_fused(u.comm, u.grad, u.n_params, s);
        if (rc) return rc;
        grad_norm_kernel<<<1, 1024, 0, s>>>(u);
        FSRL_LAUNCH_CHECK();
    }
    // torch.optim.Adam scalars (python doubles -> f32 at the op)
    const double b1 = u.beta1, b2 = u.beta2;
    const double bc1 = 1.0 - pow(b1, (double)adam_t), bc2 = 1.0 - pow(b2, (double)adam_t);
    const float neg_step = (float)(-(u.lr / bc1));
    const float bc2s = (float)sqrt(bc2);
    const int n_plain = (int)((u.n_params + 255) / 256);
    const int n_tiles = u.n_nets * (H / 32) * (H / 32);
    adam_kernel<<<n_plain + n_tiles, 256, 0, s>>>(u, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), bc2s,
                                                  (float)u.adam_eps, neg_step, slot, n_plain);
    FSRL_LAUNCH_CHECK();
    return FSRL_OK;
}

__global__ void mirror_w2_kernel(const fsrl_ppo_update_t u) {
    // w2n[n][o][k] = w2t[n][k][o]  (initial sync of the mirror, 32x32 tiles)
    __shared__ float tile[32][33];
    const int H = u.H;
    const int tpn = (H / 32) * (H / 32);
    const int n = blockIdx.x / tpn, tt = blockIdx.x % tpn;
    const int k0 = (tt / (H / 32)) * 32, o0 = (tt % (H / 32)) * 32;
    const float* src = u.theta + u.net_off[n] + (long long)u.D * H + H;
    const int lx = threadIdx.x % 32, ly = threadIdx.x / 32;
    for (int q = 0; q < 4; ++q) tile[ly + 8 * q][lx] = src[(size_t)(k0 + ly + 8 * q) * H + o0 + lx];
    __syncthreads();
    float* mir = u.w2n + (size_t)n * H * H;
    for (int q = 0; q < 4; ++q) mir[(size_t)(o0 + ly + 8 * q) * H + k0 + lx] = tile[lx][ly + 8 * q];
}

}  // namespace fsrl

using namespace fsrl;

static int check_update(const fsrl_ppo_update_t* u) {
    FSRL_REQUIRE(u != nullptr, "ppo: null descriptor");
    FSRL_REQUIRE(u->H == 64 || u->H == 128 || u->H == 256 || u->H == 512, "ppo: hidden width %d unsupported", u->H);
    FSRL_REQUIRE(u->n_nets >= 1 && u->n_nets <= 3 && u->C == u->n_nets - 1, "ppo: n_nets/C inconsistent");
    FSRL_REQUIRE(u->A >= 1 && u->A <= 8, "ppo: action dim %d out of range", u->A);
    FSRL_REQUIRE(u->theta && u->grad && u->adam_m && u->adam_v && u->w2n && u->scratch && u->norm_sq, "ppo: null buffer");
    return FSRL_OK;
}

extern "C" size_t fsrl_ppo_scratch_floats(int n_nets, int H, int bmax) {
    return (size_t)n_nets * (size_t)bmax * (4 * (size_t)H + DOUT_LD);
}

extern "C" int fsrl_ppo_sync_mirror(const fsrl_ppo_update_t* u, void* stream) {
    int rc = check_update(u);
    if (rc) return rc;
    const int H = u->H;
    mirror_w2_kernel<<<u->n_nets * (H / 32) * (H / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(*u);
    FSRL_LAUNCH_CHECK();
    return FSRL_OK;
}

// One repeat of PPOLagrangian.learn's inner loop (ppo_lag.py:223-247): every minibatch of
// Batch.split(batch_size, merge_last=True) over the permutation already in u->perm.
extern "C" int fsrl_ppo_lag_epoch(const fsrl_ppo_update_t* u, long long n_total, int batch_size,
                                  int stats_slot0, long long adam_t0, int* n_minibatches,
                                  void* stream) {
    int rc = check_update(u);
    if (rc) return rc;
    FSRL_REQUIRE(u->obs && u->act && u->logp_old && u->adv && u->ret && u->perm && u->stats, "ppo: null batch pointer");
    FSRL_REQUIRE(batch_size >= 2 && n_total >= 2, "ppo: batch too small");
    FSRL_REQUIRE(2 * batch_size - 1 <= u->bmax || n_total <= u->bmax, "ppo: scratch bmax %d too small for batch_size %d", u->bmax, batch_size);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int count = 0;
    const bool merge_last = (n_total % batch_size) > 0;       // tianshou Batch.split
    FSRL_REQUIRE(u->world <= 1 || (u->comm && u->moments_w && u->batch_size == batch_size),
                 "ppo: data-parallel run needs comm, moments buffer and batch_size in the descriptor");
    if (u->world > 1) {
        // every rank must run the same number of equally sized minibatches
        FSRL_REQUIRE(!merge_last, "ppo: data-parallel run needs n_total %% batch_size == 0");
        const int n_mb = (int)(n_total / batch_size);
        ppo_adv_moments_kernel<<<n_mb, 256, 0, s>>>(*u, n_total, n_mb);
        FSRL_LAUNCH_CHECK();
        int rc2 = fsrl_allreduce_f64(u->comm, u->moments_w, (long long)n_mb * 4, s);
        if (rc2) return rc2;
    }
    for (long long off = 0; off < n_total; off += batch_size) {
        long long B = batch_size;
        bool last = false;
        if (merge_last && off + 2LL * batch_size >= n_total) { B = n_total - off; last = true; }
        if (off + B > n_total) B = n_total - off;
        FSRL_REQUIRE(B <= u->bmax, "ppo: minibatch of %lld rows exceeds scratch (%d)", B, u->bmax);
        int r2;
        switch (u->H) {
            case 64: r2 = ppo_launch_minibatch<64>(*u, (int)off, (int)B, stats_slot0 + count, adam_t0 + count + 1, s); break;
            case 128: r2 = ppo_launch_minibatch<128>(*u, (int)off, (int)B, stats_slot0 + count, adam_t0 + count + 1, s); break;
            case 256: r2 = ppo_launch_minibatch<256>(*u, (int)off, (int)B, stats_slot0 + count, adam_t0 + count + 1, s); break;
            default: r2 = ppo_launch_minibatch<512>(*u, (int)off, (int)B, stats_slot0 + count, adam_t0 + count + 1, s); break;
        }
        if (r2) return r2;
        ++count;
        if (last) break;
    }
    if (n_minibatches) *n_minibatches = count;
    return FSRL_OK;
}

// Measurement aid for bench.py's roofline object: average duration of each phase kernel over
// `iters` back-to-back launches on one minibatch of u->perm (CUDA events on `stream`).  The
// Adam phase is launched with lr = 0 so that the weights are not disturbed.
template <int H>
static int ppo_time_phases(const fsrl_ppo_update_t& u0, int B, int iters, float* ms, cudaStream_t s) {
    using TT = MlpTile<H>;
    fsrl_ppo_update_t u = u0;
    const size_t smemA = TT::smem_bytes(u.D) + sizeof(float) * ((size_t)TT::R * TT::LDA + (size_t)TT::R * DOUT_LD);
    FSRL_CUDA(cudaFuncSetAttribute(ppo_fwdbwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    cudaEvent_t e[4];
    for (int i = 0; i < 4; ++i) FSRL_CUDA(cudaEventCreate(&e[i]));
    const dim3 gA((B + TT::R - 1) / TT::R, u.n_nets);
    const dim3 gB((H / WG_TK) * (H / WG_TO) + 2 * (H / WG_TO), u.n_nets);
    const int n_plain = (int)((u.n_params + 255) / 256);
    const int n_tiles = u.n_nets * (H / 32) * (H / 32);
    FSRL_CUDA(cudaEventRecord(e[0], s));
    for (int i = 0; i < iters; ++i) ppo_fwdbwd_kernel<H><<<gA, MLP_TPB, smemA, s>>>(u, 0, B, 0);
    FSRL_CUDA(cudaEventRecord(e[1], s));
    for (int i = 0; i < iters; ++i) ppo_wgrad_kernel<H><<<gB, WG_TPB, 0, s>>>(u, 0, B);
    FSRL_CUDA(cudaEventRecord(e[2], s));
    for (int i = 0; i < iters; ++i)
        adam_kernel<<<n_plain + n_tiles, 256, 0, s>>>(u, 0.1f, 0.999f, 0.001f, 1.0f, 1e-8f, 0.0f, -1, n_plain);
    FSRL_CUDA(cudaEventRecord(e[3], s));
    FSRL_CUDA(cudaEventSynchronize(e[3]));
    for (int i = 0; i < 3; ++i) {
        float t = 0.f;
        FSRL_CUDA(cudaEventElapsedTime(&t, e[i], e[i + 1]));
        ms[i] = t / (float)iters;
    }
    for (int i = 0; i < 4; ++i) cudaEventDestroy(e[i]);
    FSRL_LAUNCH_CHECK();
    return FSRL_OK;
}

extern "C" int fsrl_ppo_phase_times(const fsrl_ppo_update_t* u, int B, int iters, float* ms_out, void* stream) {
    int rc = check_update(u);
    if (rc) return rc;
    FSRL_REQUIRE(ms_out && iters > 0 && B > 1 && B <= u->bmax, "fsrl_ppo_phase_times: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (u->H) {
        case 64: return ppo_time_phases<64>(*u, B, iters, ms_out, s);
        case 128: return ppo_time_phases<128>(*u, B, iters, ms_out, s);
        case 256: return ppo_time_phases<256>(*u, B, iters, ms_out, s);
        default: return ppo_time_phases<512>(*u, B, iters, ms_out, s);
    }
}

extern "C" int fsrl_debug_clocks(long long* out16) {
    FSRL_CUDA(cudaMemcpyFromSymbol(out16, fsrl::g_dbg_clock, sizeof(long long) * 16));
    return FSRL_OK;
}
