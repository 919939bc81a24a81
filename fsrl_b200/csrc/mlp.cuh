// Fused 2-hidden-layer MLP forward for a tile of rows, shared by the rollout step, the critic
// evaluation before GAE, log-prob evaluation and the CPO line search.
//
// Replaces the tianshou Net/MLP/ActorProb/Critic forward the reference calls at
// fsrl/policy/base_policy.py:178 (actor) and :421-422 (critics): h = ReLU(W2 ReLU(W1 x)).
//
// Canonical parameter layout (all kernels): every Linear is stored TRANSPOSED, Wt[in][out]
// row-major, so that a warp reading one k-row of the weight touches contiguous memory
// (coalesced global loads, conflict-free LDS.128).  torch sees `.weight` as the strided view
// Wt.t().
//
// Thread mapping (256 threads): a CTA owns R = 4096/H rows; thread (tr, to) accumulates a
// 4x4 register tile (rows 4tr..4tr+3, cols 4to..4to+3).  Layer-2 weights stream through a
// cp.async double-buffered shared-memory stage of KC k-rows; activations stay in shared
// memory between layers.
#pragma once
#include "common.cuh"
#include <cuda_pipeline.h>

namespace fsrl {

struct Mlp3 {            // device pointers, canonical layout
    const float* w1t;    // [in][H]
    const float* b1;     // [H]
    const float* w2t;    // [H][H]
    const float* b2;     // [H]
    const float* w3t;    // [H][out]
    const float* b3;     // [out]
    int in, H, out;
};

constexpr int MLP_TPB = 256;
constexpr int MLP_KC = 16;          // k-rows of W2t per pipeline stage
constexpr int MLP_MAX_OUT = 16;

template <int H>
struct MlpTile {
    static_assert(H == 64 || H == 128 || H == 256 || H == 512, "hidden width must be 64/128/256/512");
    static constexpr int R = 4096 / H;          // rows per CTA
    static constexpr int TO = H / 4;            // column groups
    static constexpr int PARTS = MLP_TPB / R;   // lanes cooperating on one row in layer 3
    // shared memory (floats): x[R][in_pad] | h1[R][H] | h2[R][H] | wstage[2][KC][H]
    __host__ __device__ static constexpr int in_pad(int in) { return (in + 3) & ~3; }
    __host__ __device__ static constexpr size_t smem_bytes(int in) {
        return sizeof(float) * ((size_t)R * in_pad(in) + 2 * (size_t)R * H + 2 * (size_t)MLP_KC * H);
    }
};


// cp.async one KC x H stage of a row-major [H][H] matrix into buffer `buf` of wst
template <int H>
__device__ __forceinline__ void stage_load_hh(const float* mat, float* wst, int chunk, int buf) {
    const float* src = mat + (size_t)chunk * MLP_KC * H;
    float* dst = wst + (size_t)buf * MLP_KC * H;
    for (int i = threadIdx.x * 4; i < MLP_KC * H; i += MLP_TPB * 4)
        __pipeline_memcpy_async(dst + i, src + i, 16);
    __pipeline_commit();
}

// acc[4][4] += src[4tr+i][k] * mat[k][4to+j] over k in [0, H): `src` is a row-major [R][H]
// shared-memory tile, `mat` a row-major [H][H] global matrix streamed through `wst`.
// If stage0_in_flight the caller already issued stage_load_hh(mat, wst, 0, 0).
// All threads must call; contains __syncthreads.
template <int H>
__device__ __forceinline__ void tile_gemm_hh(float (&acc)[4][4], const float* src, const float* mat,
                                             float* wst, bool stage0_in_flight) {
    using TT = MlpTile<H>;
    const int tid = threadIdx.x;
    const int to = tid % TT::TO, tr = tid / TT::TO;
    if (!stage0_in_flight) stage_load_hh<H>(mat, wst, 0, 0);
    constexpr int NCH = H / MLP_KC;
    for (int ch = 0; ch < NCH; ++ch) {
        if (ch + 1 < NCH) stage_load_hh<H>(mat, wst, ch + 1, (ch + 1) & 1);
        if (ch + 1 < NCH) __pipeline_wait_prior(1); else __pipeline_wait_prior(0);
        __syncthreads();   // stage ch visible to all; also orders earlier smem stores of `src`
        const float* w = wst + (size_t)(ch & 1) * MLP_KC * H;
#pragma unroll
        for (int kk = 0; kk < MLP_KC; kk += 4) {
            float4 hv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                hv[i] = *reinterpret_cast<const float4*>(src + (size_t)(4 * tr + i) * H + ch * MLP_KC + kk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)(kk + q) * H + 4 * to);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x = (q == 0) ? hv[i].x : (q == 1) ? hv[i].y : (q == 2) ? hv[i].z : hv[i].w;
                    acc[i][0] = fmaf(x, wv.x, acc[i][0]); acc[i][1] = fmaf(x, wv.y, acc[i][1]);
                    acc[i][2] = fmaf(x, wv.z, acc[i][2]); acc[i][3] = fmaf(x, wv.w, acc[i][3]);
                }
            }
        }
        __syncthreads();   // everyone done with stage ch before it is overwritten
    }
}

// Computes h2 = ReLU(W2 ReLU(W1 x + b1) + b2) for the R rows already staged in xs (row-major,
// stride in_pad).  On return h2 (row-major [R][H]) is valid in shared memory for all threads.
template <int H>
__device__ __forceinline__ void mlp_hidden_forward(const Mlp3& m, const float* xs, float* h1,
                                                   float* h2, float* wst) {
    using TT = MlpTile<H>;
    const int tid = threadIdx.x;
    const int to = tid % TT::TO, tr = tid / TT::TO;
    const int inp = TT::in_pad(m.in);

    // prefetch stage 0 of W2t while layer 1 runs
    stage_load_hh<H>(m.w2t, wst, 0, 0);

    // ---- layer 1: in -> H (weights straight from L2 through the read-only path) ----------
    float acc[4][4];
    {
        const float4 b = __ldg(reinterpret_cast<const float4*>(m.b1 + 4 * to));
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w; }
        for (int k = 0; k < m.in; ++k) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(m.w1t + (size_t)k * H + 4 * to));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x = xs[(4 * tr + i) * inp + k];
                acc[i][0] = fmaf(x, w.x, acc[i][0]); acc[i][1] = fmaf(x, w.y, acc[i][1]);
                acc[i][2] = fmaf(x, w.z, acc[i][2]); acc[i][3] = fmaf(x, w.w, acc[i][3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(h1 + (size_t)(4 * tr + i) * H + 4 * to) =
                make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
    }
    // ---- layer 2: H -> H, W2t streamed through the double-buffered stage --------------------
    {
        const float4 b = __ldg(reinterpret_cast<const float4*>(m.b2 + 4 * to));
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w; }
    }
    tile_gemm_hh<H>(acc, h1, m.w2t, wst, /*stage0_in_flight=*/true);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(h2 + (size_t)(4 * tr + i) * H + 4 * to) =
            make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
    __syncthreads();
}

// Layer 3 (H -> out <= 16): PARTS lanes cooperate on each row, shuffle-reduce; on return the
// lane with part == 0 of row r (thread r*PARTS) holds out[0..out) for that row.
template <int H>
__device__ __forceinline__ void mlp_head_forward(const Mlp3& m, const float* h2, float* out) {
    using TT = MlpTile<H>;
    const int tid = threadIdx.x;
    const int r = tid / TT::PARTS, part = tid % TT::PARTS;
#pragma unroll
    for (int j = 0; j < MLP_MAX_OUT; ++j) out[j] = 0.f;
    for (int k = part; k < H; k += TT::PARTS) {
        const float x = h2[(size_t)r * H + k];
        const float* w = m.w3t + (size_t)k * m.out;
#pragma unroll
        for (int j = 0; j < MLP_MAX_OUT; ++j)
            if (j < m.out) out[j] = fmaf(x, __ldg(w + j), out[j]);
    }
#pragma unroll
    for (int j = 0; j < MLP_MAX_OUT; ++j) {
        if (j < m.out) {
            float v = out[j];
#pragma unroll
            for (int o = TT::PARTS / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, TT::PARTS);
            out[j] = v + __ldg(m.b3 + j);
        }
    }
}

}  // namespace fsrl
