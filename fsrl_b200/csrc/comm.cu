// Multi-GPU plumbing for the data-parallel update: one process per GPU, each rank
// owns its env shard and replay shard; per optimiser step ONE all-reduce of the flat gradient
// buffer over NVLink/NVSwitch (NCCL), issued from the same C loop that launches the update
// kernels so that no Python sits between the wgrad kernel, the collective and the Adam kernel.
// The reference has no distributed code at all; this is the new engine's design.
#include "common.cuh"
#include "fsrl_b200.h"
#include <nccl.h>
#include <string.h>

#define FSRL_NCCL(call)                                                                   \
    do {                                                                                  \
        ncclResult_t r__ = (call);                                                        \
        if (r__ != ncclSuccess) {                                                         \
            ::fsrl::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
            return FSRL_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

extern "C" int fsrl_comm_unique_id(char* out128) {
    FSRL_REQUIRE(out128 != nullptr, "comm: null id buffer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size changed");
    ncclUniqueId id;
    FSRL_NCCL(ncclGetUniqueId(&id));
    memcpy(out128, &id, 128);
    return FSRL_OK;
}

extern "C" int fsrl_comm_init(const char* id128, int rank, int world, void** comm_out) {
    FSRL_REQUIRE(id128 && comm_out && world >= 1 && rank >= 0 && rank < world, "comm_init: bad arguments");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t c;
    FSRL_NCCL(ncclCommInitRank(&c, world, id, rank));
    *comm_out = c;
    return FSRL_OK;
}

extern "C" int fsrl_comm_destroy(void* comm) {
    if (comm) FSRL_NCCL(ncclCommDestroy(static_cast<ncclComm_t>(comm)));
    return FSRL_OK;
}

// in-place sum over ranks of n fp32 values (gradients ++ statistics in one flat buffer)
extern "C" int fsrl_allreduce_fused(void* comm, float* buf, long long n, void* stream) {
    FSRL_REQUIRE(comm && buf && n >= 0, "allreduce: bad arguments");
    if (n == 0) return FSRL_OK;
    FSRL_NCCL(ncclAllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, static_cast<ncclComm_t>(comm),
                            static_cast<cudaStream_t>(stream)));
    return FSRL_OK;
}

extern "C" int fsrl_allreduce_f64(void* comm, double* buf, long long n, void* stream) {
    FSRL_REQUIRE(comm && buf && n >= 0, "allreduce: bad arguments");
    if (n == 0) return FSRL_OK;
    FSRL_NCCL(ncclAllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, static_cast<ncclComm_t>(comm),
                            static_cast<cudaStream_t>(stream)));
    return FSRL_OK;
}

extern "C" int fsrl_allreduce_ranges(void* comm, float* base, const long long* offs, const long long* counts,
                                     int n_ranges, void* stream) {
    FSRL_REQUIRE(comm && base && offs && counts && n_ranges >= 0, "allreduce_ranges: bad arguments");
    FSRL_NCCL(ncclGroupStart());
    for (int i = 0; i < n_ranges; ++i) {
        if (counts[i] <= 0) continue;
        FSRL_NCCL(ncclAllReduce(base + offs[i], base + offs[i], (size_t)counts[i], ncclFloat, ncclSum,
                                static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(stream)));
    }
    FSRL_NCCL(ncclGroupEnd());
    return FSRL_OK;
}

// ---- peer-memory exchange blocks (CUDA IPC; one process per GPU) ------------------------------------
extern "C" long long fsrl_p2p_stride(long long n) { return (n + 63) / 64 * 64; }
extern "C" long long fsrl_p2p_block_bytes(long long n) {
    return 2 * fsrl_p2p_stride(n) * (long long)sizeof(float) + FSRL_P2P_MAX_RANKS * 8 + 64 + FSRL_P2P_PARTIALS * (long long)sizeof(float);
}
extern "C" int fsrl_p2p_alloc(long long n_floats, void** base_out, char* ipc64_out) {
    FSRL_REQUIRE(n_floats > 0 && base_out && ipc64_out, "p2p_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size changed");
    void* p = nullptr;
    const size_t bytes = (size_t)fsrl_p2p_block_bytes(n_floats);
    FSRL_CUDA(cudaMalloc(&p, bytes));
    FSRL_CUDA(cudaMemset(p, 0, bytes));
    FSRL_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    FSRL_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(ipc64_out, &h, 64);
    *base_out = p;
    return FSRL_OK;
}
extern "C" int fsrl_p2p_open(const char* ipc64, void** peer_base_out) {
    FSRL_REQUIRE(ipc64 && peer_base_out, "p2p_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc64, 64);
    void* p = nullptr;
    FSRL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_base_out = p;
    return FSRL_OK;
}
extern "C" int fsrl_p2p_close(void* peer_base) {
    if (peer_base) FSRL_CUDA(cudaIpcCloseMemHandle(peer_base));
    return FSRL_OK;
}
extern "C" int fsrl_p2p_free(void* base) {
    if (base) FSRL_CUDA(cudaFree(base));
    return FSRL_OK;
}
extern "C" int fsrl_p2p_poll_error(const int* err_dev, int* out_host) {
    FSRL_REQUIRE(err_dev && out_host, "p2p_poll_error: null pointer");
    FSRL_CUDA(cudaMemcpy(out_host, err_dev, sizeof(int), cudaMemcpyDeviceToHost));
    return FSRL_OK;
}
