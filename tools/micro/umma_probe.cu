// Probe for the tcgen05 / bulk-copy building blocks in fsrl_b200/csrc/umma.cuh (run on a B200):
//   1. split-tf32 GEMMs from "plane layout" operand images with SWIZZLE_NONE descriptors, K-major
//      and MN-major, M = 64 (plain and lane-interleaved accumulators) and M = 128, checked against
//      an fp64 reference;
//   2. bulk-copy (cp.async.bulk) ingest bandwidth of one SM and of 96 SMs at once, L2-resident data;
//   3. latency of the device-scope flag barrier between co-resident CTAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fsrl_b200/csrc -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "umma.cuh"

using namespace fsrl::umma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct GemmCfg {
    int M, N, K;
    int a_mn, b_mn;             // 1 = MN-major
    uint32_t a_lbo, a_sbo, a_kstep, a_bytes;   // per image (hi or lo)
    uint32_t b_lbo, b_sbo, b_kstep, b_bytes;
    int interleave;             // M = 64: two N/2 MMAs, the second 16 lanes down
};

__global__ void __launch_bounds__(128) gemm_probe(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo,
                                                  GemmCfg c, float* dout, int* err) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* sa_hi = reinterpret_cast<float*>(smem);
    float* sa_lo = reinterpret_cast<float*>(smem + c.a_bytes);
    float* sb_hi = reinterpret_cast<float*>(smem + 2 * c.a_bytes);
    float* sb_lo = reinterpret_cast<float*>(smem + 2 * c.a_bytes + c.b_bytes);
    if (tid == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_done, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<64>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        mbar_expect_tx(&bar_full, 2 * c.a_bytes + 2 * c.b_bytes);
        bulk_g2s(sa_hi, a_hi, c.a_bytes, &bar_full);
        bulk_g2s(sa_lo, a_lo, c.a_bytes, &bar_full);
        bulk_g2s(sb_hi, b_hi, c.b_bytes, &bar_full);
        bulk_g2s(sb_lo, b_lo, c.b_bytes, &bar_full);
        if (!mbar_wait(&bar_full, 0)) { *err = 1; }
        tc_fence_after();
        const int nsub = c.interleave ? 2 : 1;
        const int nn = c.N / nsub;
        const uint32_t idesc = idesc_tf32(c.M, nn, c.a_mn, c.b_mn);
        for (int ks = 0; ks < c.K / 8; ++ks) {
            const uint64_t ah = smem_desc(smem_u32(sa_hi) + ks * c.a_kstep, c.a_lbo, c.a_sbo);
            const uint64_t al = smem_desc(smem_u32(sa_lo) + ks * c.a_kstep, c.a_lbo, c.a_sbo);
            for (int s = 0; s < nsub; ++s) {
                // B sub-tile s: columns [s*nn, (s+1)*nn).  K-major: rows of the plane -> + s*nn*16 bytes;
                // MN-major: columns -> (s*nn/4) planes = s*nn/4 * SBO
                const uint32_t boff = c.b_mn ? (uint32_t)(s * nn / 4) * c.b_sbo : (uint32_t)(s * nn) * 16u;
                const uint64_t bh = smem_desc(smem_u32(sb_hi) + ks * c.b_kstep + boff, c.b_lbo, c.b_sbo);
                const uint64_t bl = smem_desc(smem_u32(sb_lo) + ks * c.b_kstep + boff, c.b_lbo, c.b_sbo);
                const uint32_t d = tb + ((uint32_t)(16 * s) << 16);
                mma_tf32_ss(d, al, bh, idesc, ks > 0);
                mma_tf32_ss(d, ah, bl, idesc, true);
                mma_tf32_ss(d, ah, bh, idesc, true);
            }
        }
        mma_commit(&bar_done);
    }
    __syncwarp();
    if (!mbar_wait(&bar_done, 0)) { if ((tid & 31) == 0) *err = 2; }
    tc_fence_after();
    float v[32];
    tmem_ld32(tb + ((uint32_t)(32 * warp) << 16), v);
    for (int j = 0; j < 32; ++j) dout[(size_t)tid * 64 + j] = v[j];
    tmem_ld32(tb + ((uint32_t)(32 * warp) << 16) + 32, v);
    for (int j = 0; j < 32; ++j) dout[(size_t)tid * 64 + 32 + j] = v[j];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tb);
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static float rtf32(float x) { uint32_t b; memcpy(&b, &x, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&x, &b, 4); return x; }

// image offset (floats) of element (mn, k) of an operand whose MN extent is `MNx`, K extent `Kx`
static size_t off_kmajor(int mn, int k, int MNx) { return (size_t)(k / 4) * MNx * 4 + (size_t)mn * 4 + k % 4; }
static size_t off_mnmajor(int mn, int k, int Kx) { return (size_t)(mn / 4) * Kx * 4 + (size_t)k * 4 + mn % 4; }


// ---- accuracy study: how should the K = 256 accumulation be organised? ------------------------------------------
// nacc accumulators (K split into nacc consecutive chunks, each accumulated from zero in its own tensor-memory columns
// and summed with fp32 round-to-nearest adds afterwards); sepcorr: the two correction products (a_lo b_hi, a_hi b_lo)
// go to their own accumulators.
__global__ void __launch_bounds__(128) acc_probe(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo,
                                                 int K, int nacc, int sepcorr, float* dout, int* err) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t a_bytes = 64 * K * 4, b_bytes = 32 * K * 4;
    float* sa_hi = reinterpret_cast<float*>(smem);
    float* sa_lo = reinterpret_cast<float*>(smem + a_bytes);
    float* sb_hi = reinterpret_cast<float*>(smem + 2 * a_bytes);
    float* sb_lo = reinterpret_cast<float*>(smem + 2 * a_bytes + b_bytes);
    if (tid == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_done, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        mbar_expect_tx(&bar_full, 2 * a_bytes + 2 * b_bytes);
        bulk_g2s(sa_hi, a_hi, a_bytes, &bar_full); bulk_g2s(sa_lo, a_lo, a_bytes, &bar_full);
        bulk_g2s(sb_hi, b_hi, b_bytes, &bar_full); bulk_g2s(sb_lo, b_lo, b_bytes, &bar_full);
        if (!mbar_wait(&bar_full, 0)) *err = 1;
        tc_fence_after();
        const uint32_t idesc = idesc_tf32(64, 32, 0, 0);
        const int ksteps = K / 8, per = ksteps / nacc;
        for (int ks = 0; ks < ksteps; ++ks) {
            const int acc = ks / per;
            const bool first = (ks % per) == 0;
            const uint64_t ah = smem_desc(smem_u32(sa_hi) + ks * 2048, 1024, 128), al = smem_desc(smem_u32(sa_lo) + ks * 2048, 1024, 128);
            const uint64_t bh = smem_desc(smem_u32(sb_hi) + ks * 1024, 512, 128), bl = smem_desc(smem_u32(sb_lo) + ks * 1024, 512, 128);
            const uint32_t dmain = tb + 32 * acc, dcorr = sepcorr ? tb + 32 * (nacc + acc) : dmain;
            if (sepcorr) {
                mma_tf32_ss(dcorr, al, bh, idesc, !first);
                mma_tf32_ss(dcorr, ah, bl, idesc, true);
                mma_tf32_ss(dmain, ah, bh, idesc, !first);
            } else {
                mma_tf32_ss(dmain, al, bh, idesc, !first);
                mma_tf32_ss(dmain, ah, bl, idesc, true);
                mma_tf32_ss(dmain, ah, bh, idesc, true);
            }
        }
        mma_commit(&bar_done);
    }
    __syncwarp();
    if (!mbar_wait(&bar_done, 0)) { if ((tid & 31) == 0) *err = 2; }
    tc_fence_after();
    float sum[32], corr[32], v[32];
    for (int j = 0; j < 32; ++j) { sum[j] = 0.f; corr[j] = 0.f; }
    for (int acc = 0; acc < nacc; ++acc) {
        tmem_ld32(tb + ((uint32_t)(32 * warp) << 16) + 32 * acc, v);
        for (int j = 0; j < 32; ++j) sum[j] += v[j];
        if (sepcorr) {
            tmem_ld32(tb + ((uint32_t)(32 * warp) << 16) + 32 * (nacc + acc), v);
            for (int j = 0; j < 32; ++j) corr[j] += v[j];
        }
    }
    for (int j = 0; j < 32; ++j) dout[(size_t)tid * 32 + j] = sum[j] + corr[j];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tb);
}

static void accuracy_study() {
    const int M = 64, N = 32, K = 256;
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    // activations-like A (non-negative, as after a ReLU), weights-like B
    for (auto& x : A) { x = frand(); if (x < 0) x = 0; }
    for (auto& x : B) x = 0.1f * frand();
    std::vector<float> ah(A.size()), al(A.size()), bh(B.size()), bl(B.size());
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { float x = A[(size_t)m * K + k], h = rtf32(x), l = rtf32(x - h); size_t o = off_kmajor(m, k, M); ah[o] = h; al[o] = l; }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { float x = B[(size_t)n * K + k], h = rtf32(x), l = rtf32(x - h); size_t o = off_kmajor(n, k, N); bh[o] = h; bl[o] = l; }
    std::vector<double> ref((size_t)M * N); std::vector<float> f32((size_t)M * N);
    double rms = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double r = 0; float f = 0.f;
        for (int k = 0; k < K; ++k) { r += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k]; f = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], f); }
        ref[(size_t)m * N + n] = r; f32[(size_t)m * N + n] = f; rms += r * r;
    }
    rms = sqrt(rms / (M * N));
    double e32 = 0;
    for (size_t i = 0; i < ref.size(); ++i) e32 += (f32[i] - ref[i]) * (f32[i] - ref[i]);
    printf("[accuracy] fp32 FMA chain (CPU)             : rms err / rms ref = %.3e\n", sqrt(e32 / ref.size()) / rms);
    float *dah, *dal, *dbh, *dbl, *dd; int* derr;
    CK(cudaMalloc(&dah, A.size() * 4)); CK(cudaMalloc(&dal, A.size() * 4)); CK(cudaMalloc(&dbh, B.size() * 4)); CK(cudaMalloc(&dbl, B.size() * 4));
    CK(cudaMalloc(&dd, 128 * 32 * 4)); CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
    CK(cudaMemcpy(dah, ah.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dal, al.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbh, bh.data(), B.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbl, bl.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = 2 * A.size() * 4 + 2 * B.size() * 4;
    CK(cudaFuncSetAttribute(acc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int cfgs[6][2] = {{1, 0}, {1, 1}, {2, 0}, {4, 0}, {4, 1}, {8, 0}};
    for (auto& c : cfgs) {
        acc_probe<<<1, 128, smem>>>(dah, dal, dbh, dbl, K, c[0], c[1], dd, derr);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(128 * 32);
        CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
        double e = 0, bias = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            const int lane = (m % 16) + 32 * (m / 16);
            const double d = D[(size_t)lane * 32 + n] - ref[(size_t)m * N + n];
            e += d * d; bias += d;
        }
        printf("[accuracy] tcgen05 3xTF32, %d accumulator(s)%s : rms err / rms ref = %.3e   mean err / rms ref = %+.3e\n", c[0],
               c[1] ? ", corrections separate" : "                      ", sqrt(e / (M * N)) / rms, bias / (M * N) / rms);
    }
}

static int run_gemm(const char* name, int M, int N, int K, int a_mn, int b_mn, int interleave) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    for (auto& x : A) x = frand();
    for (auto& x : B) x = frand();
    GemmCfg c = {};
    c.M = M; c.N = N; c.K = K; c.a_mn = a_mn; c.b_mn = b_mn; c.interleave = interleave;
    c.a_bytes = (uint32_t)((size_t)M * K * 4); c.b_bytes = (uint32_t)((size_t)N * K * 4);
    if (!a_mn) { c.a_lbo = 16 * M; c.a_sbo = 128; c.a_kstep = 2 * c.a_lbo; } else { c.a_sbo = 16 * K; c.a_lbo = 128; c.a_kstep = 128; }
    if (!b_mn) { c.b_lbo = 16 * N; c.b_sbo = 128; c.b_kstep = 2 * c.b_lbo; } else { c.b_sbo = 16 * K; c.b_lbo = 128; c.b_kstep = 128; }
    std::vector<float> ah(A.size()), al(A.size()), bh(B.size()), bl(B.size());
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
        const float x = A[(size_t)m * K + k], h = rtf32(x), l = rtf32(x - h);
        const size_t o = a_mn ? off_mnmajor(m, k, K) : off_kmajor(m, k, M);
        ah[o] = h; al[o] = l;
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        const float x = B[(size_t)n * K + k], h = rtf32(x), l = rtf32(x - h);
        const size_t o = b_mn ? off_mnmajor(n, k, K) : off_kmajor(n, k, N);
        bh[o] = h; bl[o] = l;
    }
    float *dah, *dal, *dbh, *dbl, *dd; int* derr;
    CK(cudaMalloc(&dah, c.a_bytes)); CK(cudaMalloc(&dal, c.a_bytes)); CK(cudaMalloc(&dbh, c.b_bytes)); CK(cudaMalloc(&dbl, c.b_bytes));
    CK(cudaMalloc(&dd, 128 * 64 * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemset(dd, 0xff, 128 * 64 * 4)); CK(cudaMemset(derr, 0, 4));
    CK(cudaMemcpy(dah, ah.data(), c.a_bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dal, al.data(), c.a_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbh, bh.data(), c.b_bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbl, bl.data(), c.b_bytes, cudaMemcpyHostToDevice));
    const size_t smem = 2 * (size_t)c.a_bytes + 2 * (size_t)c.b_bytes;
    CK(cudaFuncSetAttribute(gemm_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_probe<<<1, 128, smem>>>(dah, dal, dbh, dbl, c, dd, derr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] KERNEL FAILED: %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * 64); int herr = 0;
    CK(cudaMemcpy(D.data(), dd, 128 * 64 * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0.0, maxref = 0.0; int bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0.0;
        for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
        int lane, col;
        if (M == 128) { lane = m; col = n; }
        else {
            lane = (m % 16) + 32 * (m / 16); col = n;
            if (interleave && n >= N / 2) { lane += 16; col = n - N / 2; }
        }
        const double got = D[(size_t)lane * 64 + col];
        const double er = fabs(got - ref);
        if (!(er <= 1e-4)) ++bad;
        if (er > maxerr || er != er) maxerr = er;
        if (fabs(ref) > maxref) maxref = fabs(ref);
    }
    printf("[%s] M=%d N=%d K=%d a_mn=%d b_mn=%d il=%d : err_flag=%d max|err|=%.3e (max|ref|=%.2f) bad=%d/%d  -> %s\n", name, M, N, K, a_mn,
           b_mn, interleave, herr, maxerr, maxref, bad, M * N, (bad == 0 && herr == 0) ? "OK" : "MISMATCH");
    if (bad) {   // show where the first rows went
        for (int lane = 0; lane < 128; lane += 8) printf("   lane %3d: %.4f %.4f %.4f\n", lane, D[(size_t)lane * 64], D[(size_t)lane * 64 + 1], D[(size_t)lane * 64 + 16]);
        double r00 = 0, r10 = 0, r01 = 0;
        for (int k = 0; k < K; ++k) { r00 += (double)A[k] * B[k]; r10 += (double)A[(size_t)K + k] * B[k]; r01 += (double)A[k] * B[(size_t)K + k]; }
        printf("   ref D[0][0]=%.4f D[1][0]=%.4f D[0][1]=%.4f\n", r00, r10, r01);
    }
    cudaFree(dah); cudaFree(dal); cudaFree(dbh); cudaFree(dbl); cudaFree(dd); cudaFree(derr);
    return bad != 0;
}

// ---- ingest bandwidth ------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ingest_probe(const float* src, size_t region_floats, int share, int iters, int chunk_bytes,
                                                    int nchunks, long long* cycles, int* err) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[8];
    const int tid = threadIdx.x;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    __syncthreads();
    const float* mine = src + (size_t)(blockIdx.x / share) * region_floats;
    if (tid == 0) {
        long long best = 1LL << 60, tot = 0;
        for (int it = 0; it < iters; ++it) {
            const long long t0 = clock64();
            for (int ch = 0; ch < nchunks; ++ch) {
                mbar_expect_tx(&bar[ch], chunk_bytes);
                bulk_g2s(smem + (size_t)ch * chunk_bytes, reinterpret_cast<const unsigned char*>(mine) + (size_t)ch * chunk_bytes, chunk_bytes, &bar[ch]);
            }
            for (int ch = 0; ch < nchunks; ++ch) if (!mbar_wait(&bar[ch], it & 1)) *err = 3;
            const long long dt = clock64() - t0;
            if (it > 0) { tot += dt; if (dt < best) best = dt; }
        }
        cycles[2 * blockIdx.x] = best; cycles[2 * blockIdx.x + 1] = tot / (iters - 1);
    }
}

__global__ void __launch_bounds__(128) barrier_probe(unsigned* ctr, int iters, long long* cycles, int* err) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            __threadfence();
            flag_add_release(ctr);
            if (!flag_wait_ge(ctr, (unsigned)(it + 1) * gridDim.x)) { *err = 4; break; }
        }
        cycles[blockIdx.x] = (clock64() - t0) / iters;
    }
}

int main() {
    srand(1234);
    int fails = 0;
    fails += run_gemm("G1  K-major x K-major", 64, 32, 256, 0, 0, 0);
    fails += run_gemm("G1i K-major x K-major interleaved", 64, 32, 256, 0, 0, 1);
    fails += run_gemm("G3  MN x MN", 64, 64, 128, 1, 1, 0);
    fails += run_gemm("G3i MN x MN interleaved", 64, 64, 128, 1, 1, 1);
    fails += run_gemm("G2  MN x K", 64, 64, 32, 1, 0, 0);
    fails += run_gemm("G2i MN x K interleaved", 64, 64, 32, 1, 0, 1);
    fails += run_gemm("M128 K x K", 128, 32, 128, 0, 0, 0);
    fails += run_gemm("M128 MN x MN", 128, 64, 64, 1, 1, 0);

    accuracy_study();
    // ingest: every CTA pulls 3 x 64 KB per iteration from an L2-resident region
    int* derr; CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
    const int chunk = 64 * 1024, nch = 3;
    const size_t region = (size_t)chunk * nch / 4;
    float* src; CK(cudaMalloc(&src, region * 4 * 148)); CK(cudaMemset(src, 0, region * 4 * 148));
    long long* dcy; CK(cudaMalloc(&dcy, 2 * 148 * sizeof(long long)));
    CK(cudaFuncSetAttribute(ingest_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk * nch));
    const int grids[4] = {1, 32, 96, 144};
    for (int share = 1; share <= 8; share *= 8) {
        for (int gi = 0; gi < 4; ++gi) {
            const int g = grids[gi];
            ingest_probe<<<g, 128, chunk * nch>>>(src, region, share, 20, chunk, nch, dcy, derr);
            CK(cudaDeviceSynchronize());
            std::vector<long long> cy(2 * g);
            CK(cudaMemcpy(cy.data(), dcy, 2 * g * sizeof(long long), cudaMemcpyDeviceToHost));
            long long bmin = 1LL << 60, bmax = 0; double avg = 0;
            for (int i = 0; i < g; ++i) { bmin = cy[2 * i] < bmin ? cy[2 * i] : bmin; bmax = cy[2 * i] > bmax ? cy[2 * i] : bmax; avg += cy[2 * i + 1]; }
            avg /= g;
            printf("[ingest] grid=%3d share=%d : 192 KB per CTA: best %lld..%lld cycles, mean %.0f  -> %.1f B/clk/SM (mean)\n", g, share, bmin, bmax, avg,
                   (double)chunk * nch / avg);
        }
    }
    // small-chunk variant: 24 x 8 KB
    CK(cudaFuncSetAttribute(ingest_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 24 * 1024));
    // (only 8 barriers: use 8 chunks of 24 KB)
    ingest_probe<<<96, 128, 8 * 24 * 1024>>>(src, region, 1, 20, 24 * 1024, 8, dcy, derr);
    CK(cudaDeviceSynchronize());
    {
        std::vector<long long> cy(2 * 96);
        CK(cudaMemcpy(cy.data(), dcy, 2 * 96 * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 96; ++i) avg += cy[2 * i + 1]; avg /= 96;
        printf("[ingest] grid= 96 8 x 24 KB: mean %.0f cycles -> %.1f B/clk/SM\n", avg, 8.0 * 24 * 1024 / avg);
    }
    unsigned* ctr; CK(cudaMalloc(&ctr, 4));
    for (int gi = 0; gi < 4; ++gi) {
        const int g = grids[gi] == 1 ? 8 : grids[gi];
        CK(cudaMemset(ctr, 0, 4));
        barrier_probe<<<g, 128>>>(ctr, 200, dcy, derr);
        CK(cudaDeviceSynchronize());
        std::vector<long long> cy(g);
        CK(cudaMemcpy(cy.data(), dcy, g * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < g; ++i) avg += cy[i]; avg /= g;
        printf("[barrier] %3d CTAs: %.0f cycles per barrier\n", g, avg);
    }
    int herr = 0; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("err flag %d, sm clock attr %d kHz, gemm fails %d\n", herr, clk, fails);
    return 0;
}
