// Dual GAE(lambda): reward AND cost advantages/returns in ONE pass over the flattened,
// env-major (env, step) buffer.
//
// Replaces (reference, CPU):
//   /root/reference/fsrl/policy/base_policy.py:524-540  gae_return (numba, sequential, f64)
//   /root/reference/fsrl/policy/base_policy.py:409-411,429,438-446  value_mask / end_flag /
//       ret = adv + v / cast to f32, for both critics
//
// Algorithm.  The recurrence  g_i = delta_i + a_i * g_{i+1},  a_i = (1-end_i)*gamma*lambda,
// is a reverse scan of affine maps  f_i(g) = b_i + a_i*g  under composition
//   (A1,B1) o (A2,B2) = (A1*A2, B1 + A1*B2).
// Segment ends need no special casing: a_i = 0 annihilates the carry.  Both critics share
// a_i, so the scan state is (A, B_rew, B_cost) in f64 (the reference accumulates in f64).
//
// Single pass, HBM-bound: tiles of TILE elements are claimed in reverse memory order via an
// atomic ticket; each thread runs the recurrence sequentially over ITEMS contiguous
// elements (same op order as the reference inside a thread), thread aggregates are combined
// with a warp-shuffle scan + one smem hop, and the carry across tiles uses decoupled
// look-back (aggregate / inclusive-prefix descriptors).  Algorithmic traffic:
// 16*C + 10 = 42 B per transition for C = 2.
#include "common.cuh"

namespace fsrl {

constexpr int GAE_TPB = 256;
constexpr int GAE_ITEMS = 8;
constexpr int GAE_TILE = GAE_TPB * GAE_ITEMS;  // 2048 transitions per tile

struct __align__(16) GaeState {  // affine map (A, B[2])
    double A, Br, Bc;
};

__device__ __forceinline__ GaeState compose(const GaeState& later, const GaeState& earlier) {
    // "later" is applied after "earlier" in scan order (scan order == reverse memory order)
    GaeState r;
    r.A = later.A * earlier.A;
    r.Br = later.Br + later.A * earlier.Br;
    r.Bc = later.Bc + later.A * earlier.Bc;
    return r;
}

__device__ __forceinline__ GaeState shfl_up(const GaeState& s, int d) {
    GaeState r;
    r.A = __shfl_up_sync(0xffffffffu, s.A, d);
    r.Br = __shfl_up_sync(0xffffffffu, s.Br, d);
    r.Bc = __shfl_up_sync(0xffffffffu, s.Bc, d);
    return r;
}

// Tile descriptor for decoupled look-back.  status: 0 = empty, 1 = aggregate, 2 = inclusive.
struct GaeDesc {
    double A, Br, Bc;  // aggregate of this tile
    double Ir, Ic;     // inclusive carry leaving this tile (the g value at its first element)
    int status;
    int pad;
};

struct GaeWorkspace {
    unsigned int ticket;
    unsigned int pad[3];
    // followed by GaeDesc[num_tiles]
};

template <int C, bool VEC>
__global__ void __launch_bounds__(GAE_TPB, 3)
gae_dual_kernel(const float* __restrict__ v, const float* __restrict__ vnext,
                const float* __restrict__ rew, const float* __restrict__ cost,
                const uint8_t* __restrict__ end_flag, const uint8_t* __restrict__ terminated,
                double gamma, double gl, float* __restrict__ adv, float* __restrict__ ret,
                long long N, long long ld, int num_tiles, GaeWorkspace* ws) {
    GaeDesc* desc = reinterpret_cast<GaeDesc*>(ws + 1);
    __shared__ unsigned int s_ticket;
    __shared__ GaeState s_warp[GAE_TPB / 32];
    __shared__ double s_carry[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;

    // persistent CTAs: the grid is one resident wave (or fewer); every CTA keeps claiming tiles in
    // scan order until the ticket counter runs out, so there is no partial second wave.  Look-back
    // only ever waits on smaller tickets, which were claimed earlier by CTAs that are running.
    for (;;) {
    __syncthreads();                           // s_ticket / s_warp / s_carry of the previous tile are dead
    if (tid == 0) s_ticket = atomicAdd(&ws->ticket, 1u);
    __syncthreads();
    const int ticket = (int)s_ticket;          // scan-order tile index (0 = end of memory)
    if (ticket >= num_tiles) break;
    const int m = num_tiles - 1 - ticket;      // memory tile index
    // thread t owns scan positions [t*ITEMS, (t+1)*ITEMS) of the tile == memory chunk
    // starting at base, walked backwards
    const long long base = (long long)m * GAE_TILE + (long long)(GAE_TPB - 1 - tid) * GAE_ITEMS;

    float fv[C][GAE_ITEMS], fvn[C][GAE_ITEMS], fm[C][GAE_ITEMS];
    uint8_t fe[GAE_ITEMS], ft[GAE_ITEMS];
    const bool full = base + GAE_ITEMS <= N;
    if (VEC && full) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float* pm = (c == 0) ? rew : cost;
#pragma unroll
            for (int q = 0; q < GAE_ITEMS / 4; ++q) {
                float4 a = ldg_stream4(v + c * ld + base + 4 * q);
                float4 b = ldg_stream4(vnext + c * ld + base + 4 * q);
                float4 d = ldg_stream4(pm + base + 4 * q);
                fv[c][4 * q] = a.x; fv[c][4 * q + 1] = a.y; fv[c][4 * q + 2] = a.z; fv[c][4 * q + 3] = a.w;
                fvn[c][4 * q] = b.x; fvn[c][4 * q + 1] = b.y; fvn[c][4 * q + 2] = b.z; fvn[c][4 * q + 3] = b.w;
                fm[c][4 * q] = d.x; fm[c][4 * q + 1] = d.y; fm[c][4 * q + 2] = d.z; fm[c][4 * q + 3] = d.w;
            }
        }
        static_assert(GAE_ITEMS == 8, "flag loads assume 8 items");
        uint2 e8 = __ldcs(reinterpret_cast<const uint2*>(end_flag + base));
        uint2 t8 = terminated ? __ldcs(reinterpret_cast<const uint2*>(terminated + base)) : make_uint2(0, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            fe[j] = (e8.x >> (8 * j)) & 0xff; fe[4 + j] = (e8.y >> (8 * j)) & 0xff;
            ft[j] = (t8.x >> (8 * j)) & 0xff; ft[4 + j] = (t8.y >> (8 * j)) & 0xff;
        }
    } else {
#pragma unroll
        for (int j = 0; j < GAE_ITEMS; ++j) {
            const long long i = base + j;
            const bool ok = i < N;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float* pm = (c == 0) ? rew : cost;
                fv[c][j] = ok ? v[c * ld + i] : 0.f;
                fvn[c][j] = ok ? vnext[c * ld + i] : 0.f;
                fm[c][j] = ok ? pm[i] : 0.f;
            }
            // out-of-range padding behaves as an identity element AFTER the data in memory
            // (it is scanned first): a = 0, delta = 0 -> contributes g = 0 and kills nothing.
            fe[j] = ok ? end_flag[i] : 1;
            ft[j] = (ok && terminated) ? terminated[i] : 0;
        }
    }

    // ---- pass 1: per-thread aggregate with zero carry-in (scan order = j descending) ----
    // delta_j and a_j are recomputed in pass 2 (3 f64 ops each, bit-identical) instead of being kept
    // in 48 registers: one more resident CTA per SM matters more to this HBM-bound kernel
    auto delta = [&](int c, int j) {
        // value_mask (:429): v_next * ~terminated, then delta = rew + v_next*gamma - v (:534)
        const double vn = ft[j] ? 0.0 : (double)fvn[c][j];
        return __dsub_rn(__dadd_rn((double)fm[c][j], __dmul_rn(vn, gamma)), (double)fv[c][j]);
    };
    GaeState agg;
    agg.A = 1.0; agg.Br = 0.0; agg.Bc = 0.0;
#pragma unroll
    for (int j = GAE_ITEMS - 1; j >= 0; --j) {
        const double aj = fe[j] ? 0.0 : gl;   // (1.0 - end) * (gamma*lambda): exactly 0 or gl
        agg.Br = __dadd_rn(delta(0, j), __dmul_rn(aj, agg.Br));
        if (C > 1) agg.Bc = __dadd_rn(delta(C - 1, j), __dmul_rn(aj, agg.Bc));
        agg.A *= aj;
    }

    // ---- block scan of thread aggregates (inclusive, scan order = tid ascending) ---------
    GaeState inc = agg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        GaeState o = shfl_up(inc, d);
        if (lane >= d) inc = compose(inc, o);
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    // exclusive prefix over preceding warps
    GaeState wpre; wpre.A = 1.0; wpre.Br = 0.0; wpre.Bc = 0.0;
    for (int w = 0; w < wid; ++w) wpre = compose(s_warp[w], wpre);
    GaeState ex = shfl_up(inc, 1);            // exclusive within warp
    if (lane == 0) { ex.A = 1.0; ex.Br = 0.0; ex.Bc = 0.0; }
    ex = compose(ex, wpre);                   // exclusive prefix within the tile
    // tile aggregate = inclusive of the last thread
    // ---- decoupled look-back for the carry entering this tile ------------------------------
    if (tid == GAE_TPB - 1) {
        GaeState tile = compose(inc, wpre);
        double cr = 0.0, cc = 0.0;
        if (ticket == 0) {
            desc[0].Ir = tile.Br; desc[0].Ic = tile.Bc;
            __threadfence();
            atomicExch(&desc[0].status, 2);
        } else {
            desc[ticket].A = tile.A; desc[ticket].Br = tile.Br; desc[ticket].Bc = tile.Bc;
            __threadfence();
            atomicExch(&desc[ticket].status, 1);
            // walk predecessors (smaller ticket = earlier in scan order)
            GaeState acc; acc.A = 1.0; acc.Br = 0.0; acc.Bc = 0.0;  // composition of tiles (ticket-1 .. p+1)
            int p = ticket - 1;
            while (true) {
                int st;
                do { st = atomicAdd(&desc[p].status, 0); } while (st == 0);
                __threadfence();
                if (st == 2) {
                    const double ir = __ldcg(&desc[p].Ir), ic = __ldcg(&desc[p].Ic);
                    cr = acc.Br + acc.A * ir;
                    cc = acc.Bc + acc.A * ic;
                    break;
                }
                GaeState t; t.A = __ldcg(&desc[p].A); t.Br = __ldcg(&desc[p].Br); t.Bc = __ldcg(&desc[p].Bc);
                acc = compose(acc, t);
                if (acc.A == 0.0 || p == 0) {   // a segment end inside: nothing older matters
                    cr = acc.Br; cc = acc.Bc;
                    if (p == 0 && acc.A != 0.0) { cr = acc.Br; cc = acc.Bc; }
                    break;
                }
                --p;
            }
            desc[ticket].Ir = tile.Br + tile.A * cr;
            desc[ticket].Ic = tile.Bc + tile.A * cc;
            __threadfence();
            atomicExch(&desc[ticket].status, 2);
        }
        s_carry[0] = cr; s_carry[1] = cc;
    }
    __syncthreads();
    // carry entering this thread's chunk
    double gr = ex.Br + ex.A * s_carry[0];
    double gc = ex.Bc + ex.A * s_carry[1];

    // ---- pass 2: re-run the recurrence with the true carry, emit adv / ret ------------------
    float oa[C][GAE_ITEMS], orr[C][GAE_ITEMS];
#pragma unroll
    for (int j = GAE_ITEMS - 1; j >= 0; --j) {
        const double aj = fe[j] ? 0.0 : gl;
        gr = __dadd_rn(delta(0, j), __dmul_rn(aj, gr));
        oa[0][j] = (float)gr;
        orr[0][j] = (float)__dadd_rn(gr, (double)fv[0][j]);      // ret = adv + v (:441)
        if (C > 1) {
            gc = __dadd_rn(delta(C - 1, j), __dmul_rn(aj, gc));
            oa[C - 1][j] = (float)gc;
            orr[C - 1][j] = (float)__dadd_rn(gc, (double)fv[C - 1][j]);
        }
    }
    if (VEC && full) {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int q = 0; q < GAE_ITEMS / 4; ++q) {
                stg_stream4(adv + c * ld + base + 4 * q,
                            make_float4(oa[c][4 * q], oa[c][4 * q + 1], oa[c][4 * q + 2], oa[c][4 * q + 3]));
                stg_stream4(ret + c * ld + base + 4 * q,
                            make_float4(orr[c][4 * q], orr[c][4 * q + 1], orr[c][4 * q + 2], orr[c][4 * q + 3]));
            }
    } else {
#pragma unroll
        for (int j = 0; j < GAE_ITEMS; ++j) {
            const long long i = base + j;
            if (i < N) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    adv[c * ld + i] = oa[c][j];
                    ret[c * ld + i] = orr[c][j];
                }
            }
        }
    }
    }   // next ticket
}

}  // namespace fsrl

using namespace fsrl;

extern "C" size_t fsrl_gae_dual_workspace_bytes(int64_t N) {
    const int64_t tiles = (N + GAE_TILE - 1) / GAE_TILE;
    return sizeof(GaeWorkspace) + (size_t)(tiles > 0 ? tiles : 1) * sizeof(GaeDesc);
}

extern "C" int fsrl_gae_dual(const float* v, const float* vnext, const float* rew,
                             const float* cost, const uint8_t* end_flag,
                             const uint8_t* terminated, double gamma, double gae_lambda,
                             float* adv, float* ret, int64_t N, int64_t ld, int C,
                             void* workspace, size_t workspace_bytes, void* stream) {
    FSRL_REQUIRE(N >= 0, "fsrl_gae_dual: N must be >= 0 (got %lld)", (long long)N);
    FSRL_REQUIRE(C == 1 || C == 2, "fsrl_gae_dual: C must be 1 or 2 (got %d)", C);
    FSRL_REQUIRE(gamma >= 0.0 && gamma <= 1.0, "discount factor should be in [0, 1].");
    FSRL_REQUIRE(gae_lambda >= 0.0 && gae_lambda <= 1.0, "GAE lambda should be in [0, 1].");
    if (N == 0) return FSRL_OK;
    FSRL_REQUIRE(v && vnext && rew && end_flag && adv && ret, "fsrl_gae_dual: null pointer");
    FSRL_REQUIRE(C == 1 || cost, "fsrl_gae_dual: cost pointer required for C == 2");
    FSRL_REQUIRE(ld >= N, "fsrl_gae_dual: ld (%lld) < N (%lld)", (long long)ld, (long long)N);
    const size_t need = fsrl_gae_dual_workspace_bytes(N);
    if (!workspace || workspace_bytes < need) {
        set_error("fsrl_gae_dual: workspace too small (%zu < %zu)", workspace_bytes, need);
        return FSRL_EWORKSPACE;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int tiles = (int)((N + GAE_TILE - 1) / GAE_TILE);
    FSRL_CUDA(cudaMemsetAsync(workspace, 0, need, s));
    const bool vec = aligned16(v) && aligned16(vnext) && aligned16(rew) && aligned16(adv) &&
                     aligned16(ret) && (C == 1 || aligned16(cost)) && (ld % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(end_flag) & 7u) == 0) &&
                     (!terminated || (reinterpret_cast<uintptr_t>(terminated) & 7u) == 0);
    const double gl = gamma * gae_lambda;
    GaeWorkspace* ws = static_cast<GaeWorkspace*>(workspace);
    const int wave = 3 * sm_count();                   // __launch_bounds__(GAE_TPB, 3)
    const int grid = tiles < wave ? tiles : wave;
#define LAUNCH(CC, VV)                                                                       \
    gae_dual_kernel<CC, VV><<<grid, GAE_TPB, 0, s>>>(v, vnext, rew, cost, end_flag, terminated, \
                                                     gamma, gl, adv, ret, (long long)N,      \
                                                     (long long)ld, tiles, ws)
    if (C == 2) { if (vec) LAUNCH(2, true); else LAUNCH(2, false); }
    else        { if (vec) LAUNCH(1, true); else LAUNCH(1, false); }
#undef LAUNCH
    FSRL_LAUNCH_CHECK();
    return FSRL_OK;
}
