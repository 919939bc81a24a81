// Issue-rate microbenchmark of the legacy warp-level tensor path on sm_100a:
// independent mma.sync chains per warp, W warps per SM sub-partition; prints cycles per MMA per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void tf32_kernel(int iters, long long* out, float* sink) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f900000u, 0x3fa00000u};
    uint32_t b[2] = {0x3f800000u, 0x3f810000u + threadIdx.x};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CHAINS>
__global__ void bf16_kernel(int iters, long long* out, float* sink) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f903f80u, 0x3fa03f80u};
    uint32_t b[2] = {0x3f803f80u, 0x3f813f80u + threadIdx.x};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CHAINS>
__global__ void ffma_kernel(int iters, long long* out, float* sink) {
    float c[CHAINS];
    for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x * 0.001f;
    float a = 1.0001f + threadIdx.x * 1e-6f, b = 0.9999f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[i]) : "f"(a), "f"(b));
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    long long* d; float* sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
    const int iters = 2000;
    auto report = [&](const char* name, int warps, int chains, long long macs_per_instr) {
        long long cyc; cudaDeviceSynchronize(); cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const double per_smsp_instr = (double)iters * chains * (warps / 4.0 > 1 ? warps / 4.0 : 1.0);
        printf("%-6s warps/CTA %2d chains %2d : %8lld cycles  -> %.2f cycles/instr/SMSP, %.0f MAC/clk/SM\n", name, warps,
               chains, cyc, cyc / per_smsp_instr, (double)iters * chains * warps * macs_per_instr / cyc);
    };
    for (int warps : {4, 8, 16}) {
        tf32_kernel<1><<<148, warps * 32>>>(iters, d, sink); report("tf32", warps, 1, 1024);
        tf32_kernel<4><<<148, warps * 32>>>(iters, d, sink); report("tf32", warps, 4, 1024);
        tf32_kernel<8><<<148, warps * 32>>>(iters, d, sink); report("tf32", warps, 8, 1024);
        bf16_kernel<1><<<148, warps * 32>>>(iters, d, sink); report("bf16", warps, 1, 2048);
        bf16_kernel<8><<<148, warps * 32>>>(iters, d, sink); report("bf16", warps, 8, 2048);
        ffma_kernel<8><<<148, warps * 32>>>(iters, d, sink); report("ffma", warps, 8, 32);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
