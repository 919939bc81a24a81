// sm_100a building blocks for the persistent PPO update (csrc/ppo_persist.cu): tcgen05 tensor
// core MMAs (kind::tf32, accumulators in tensor memory), bulk asynchronous copies (TMA unit,
// cp.async.bulk) completing on mbarriers, and the device-scope flag barriers that chain the
// CTAs of the persistent grid.  Inline PTX only -- no CUTLASS / CuTe dependency.
//
// Operand layout ("plane layout", no swizzle).  A matrix X[mn][k] of fp32 words is stored as
//       X_img[k / 4][mn][k % 4]            (planes of R rows x 16 bytes, R = MN extent of the block)
// which is the canonical K-major SWIZZLE_NONE layout of tcgen05.mma (CUTLASS cute/atom/mma_traits_sm100.hpp,
// make_umma_desc, "INTERLEAVE" rows): 8 x 16 B core matrices are 8 consecutive rows of one plane,
//         SBO (next 8 rows) = 128 B,   LBO (next 4 k = next plane) = 16 * R bytes,
// and one MMA instruction consumes K = 8 tf32 values = 2 planes.  tf32 operands can be read MN-major only from
// the 128B_BASE32B swizzled layout (measured: tools/micro/umma_probe.cu -- the SWIZZLE_NONE MN-major
// descriptor silently yields zeros), so an operand that is contracted over its other index is published by its
// producer as a second, transposed K-major image.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fsrl {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a barrier that never completes must not hang the GPU (returns false on timeout)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, long long timeout_cycles = 4000000000LL) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > timeout_cycles) return false;
    }
    return true;
}

// ---- bulk asynchronous copy global -> shared (TMA unit; SASS UBLKCP), completes on an mbarrier --
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// orders this thread's earlier generic-proxy operations (the acquire of a flag) before its later
// async-proxy operations (bulk copies reading what another SM wrote with ordinary stores)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- tensor memory ------------------------------------------------------------------------------
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {        // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive 32-bit columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- tcgen05.mma kind::tf32, operands from shared memory ------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor: start >> 4 at [0,14),
// LBO >> 4 at [16,30), SBO >> 4 at [32,46), version 1 at [46,48), layout_type 0 at [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32
constexpr uint32_t IDESC_MN_MAJOR_A = 1u << 15, IDESC_MN_MAJOR_B = 1u << 16;
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major ? IDESC_MN_MAJOR_A : 0u) | (b_mn_major ? IDESC_MN_MAJOR_B : 0u) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// all MMAs issued so far by this thread arrive (once) on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- device-scope flag barriers between the CTAs of a co-resident grid ---------------------------
__device__ __forceinline__ void flag_add_release(unsigned* ctr, unsigned v = 1u) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned flag_ld_acquire(const unsigned* ctr) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    return v;
}
// spin until *ctr >= target; false after `timeout_cycles` (caller raises the error flag and leaves).  BACKOFF: sleep
// ~20 ns between polls.  Measured on the persistent PPO kernel: polling flat out costs 3k cycles per minibatch step
// (53.2k vs 50.0k) -- the pollers compete with the warps that share their scheduler and with the flag's L2 slice.
template <bool BACKOFF = true>
__device__ __forceinline__ bool flag_wait_ge(const unsigned* ctr, unsigned target, long long timeout_cycles = 4000000000LL) {
    if (flag_ld_acquire(ctr) >= target) return true;
    const long long t0 = clock64();
    while (flag_ld_acquire(ctr) < target) {
        if (clock64() - t0 > timeout_cycles) return false;
        if (BACKOFF) __nanosleep(20);
    }
    return true;
}

// fp32 -> (hi, lo) tf32 pair with the 13 low mantissa bits cleared (round to nearest, ties away):
// x ~= hi + lo to ~2^-22 relative; whatever the tensor core does with the low bits is irrelevant
__device__ __forceinline__ float tf32_round(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
    hi = tf32_round(x);
    lo = tf32_round(x - hi);
}

}  // namespace umma
}  // namespace fsrl
