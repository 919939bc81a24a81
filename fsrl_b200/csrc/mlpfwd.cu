// Batched MLP forward over N rows (optionally gathered by index): the critic passes of
// compute_gae_returns (/root/reference/fsrl/policy/base_policy.py:416-422) and any other
// "no_grad forward over the whole buffer" of the reference (ppo_lag.py:144-149, cpo.py:135-141).
#include "mlp.cuh"
#include "fsrl_b200.h"

namespace fsrl {

template <int H>
__global__ void __launch_bounds__(MLP_TPB)
mlp_forward_kernel(const Mlp3 m, const float* __restrict__ x, const int* __restrict__ idx,
                   long long n_rows, float* __restrict__ y) {
    using TT = MlpTile<H>;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const MlpSmem<H> sm(smem, m.in, m.out);
    const long long r0 = (long long)blockIdx.x * TT::R;
    mlp_stage_rows<H>(sm, m.in, [&](int r) -> const float* {
        const long long row = r0 + r;
        if (row >= n_rows) return nullptr;
        const long long src = idx ? (long long)idx[row] : row;
        return x + src * m.in;
    });
    __syncthreads();
    mlp_hidden_forward<H>(m, sm);
    float out[MLP_MAX_OUT];
    mlp_head_forward<H>(m, sm, out);
    const int r = tid / TT::PARTS, part = tid % TT::PARTS;
    const long long row = r0 + r;
    if (part == 0 && row < n_rows) {
#pragma unroll
        for (int j = 0; j < MLP_MAX_OUT; ++j)
            if (j < m.out) y[row * m.out + j] = out[j];
    }
}

}  // namespace fsrl

using namespace fsrl;

extern "C" int fsrl_mlp_forward(const fsrl_mlp3_t* net, const float* x, const int* idx,
                                long long n_rows, float* y, void* stream) {
    FSRL_REQUIRE(n_rows >= 0, "fsrl_mlp_forward: n_rows < 0");
    if (n_rows == 0) return FSRL_OK;
    FSRL_REQUIRE(net && x && y, "fsrl_mlp_forward: null pointer");
    FSRL_REQUIRE(net->out >= 1 && net->out <= MLP_MAX_OUT, "fsrl_mlp_forward: out dim %d unsupported", net->out);
    if (n_rows == 0) return FSRL_OK;
    const Mlp3 m = *reinterpret_cast<const Mlp3*>(net);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define GO(HH)                                                                                     \
    {                                                                                              \
        using TT = MlpTile<HH>;                                                                    \
        const size_t smem = TT::smem_bytes(m.in);                                                  \
        FSRL_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        const long long grid = (n_rows + TT::R - 1) / TT::R;                                       \
        mlp_forward_kernel<HH><<<(unsigned)grid, MLP_TPB, smem, s>>>(m, x, idx, n_rows, y);        \
    }
    switch (m.H) {
        case 64: GO(64) break;
        case 128: GO(128) break;
        case 256: GO(256) break;
        case 512: GO(512) break;
        default: set_error("fsrl_mlp_forward: hidden width %d unsupported (64/128/256/512)", m.H); return FSRL_EINVAL;
    }
#undef GO
    FSRL_LAUNCH_CHECK();
    return FSRL_OK;
}
