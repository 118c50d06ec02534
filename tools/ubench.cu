// Micro-benchmarks that calibrate the attention kernel design on B200 (run via tools/run_ubench.sh on the GPU box):
// tcgen05.ld / tcgen05.st throughput per SM, MUFU ex2 throughput, tcgen05.mma rate by shape and A-operand source.
// Not part of the product: nothing in perceptor_b200/ links this.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../perceptor_b200/csrc/pcg_ptx.cuh"

using namespace pcg;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode 0: tcgen05.ld, 1: tcgen05.st, 2: ex2 only, 3: ld + ex2 on every element + st of packed bf16 (softmax-like)
__global__ void __launch_bounds__(512, 1) k_tmem(int mode, int iters, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        tmem_alloc(&slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = threadIdx.x + j;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
            tmem_ld_32x32(trow + ((i * 32 + (warp >> 2) * 64) & 511 & ~31), v);
            if ((i & 3) == 3) tmem_wait_ld();
        }
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j]);
    } else if (mode == 1) {
        for (int i = 0; i < iters; ++i) tmem_st_32x32(trow + ((i * 32) & 511 & ~31), v);
        tmem_wait_st();
    } else if (mode == 2) {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = -1.0f - 0.001f * j - 0.0001f * threadIdx.x;
        for (int i = 0; i < iters * 4; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = exp2f(x[j]) - 1.5f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += x[j];
    } else {
        for (int i = 0; i < iters; ++i) {
            const uint32_t col = (i * 32) & 255;
            tmem_ld_32x32(trow + col, v);
            tmem_wait_ld();
            float e[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                e[j] = exp2f(fmaf(__uint_as_float(v[j]), 1.44f, -3.0f));
                acc += e[j];
            }
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack_bf16(e[2 * j], e[2 * j + 1]);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(trow + 256 + (col >> 1)),
                "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
                "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
                : "memory");
        }
        tmem_wait_st();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// tcgen05.mma rate: one thread issues `n_mma` MMAs (M=128, N=n, K=16) back to back, A from smem (ts=0) or TMEM (ts=1)
__global__ void __launch_bounds__(128, 1) k_mma(int n, int ts, int b_mn, int n_mma, int nacc, long long* cycles) {
    extern __shared__ uint8_t raw[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const uint32_t addr = smem_u32(raw);
    uint8_t* base = raw + (((addr + 1023u) & ~1023u) - addr);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        tmem_alloc(&slot, 512);
        tmem_relinquish();
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, n, 0, b_mn);
        const uint64_t da = umma_smem_desc_sw128(smem_u32(base));
        const uint64_t db = umma_smem_desc_sw128(smem_u32(base + 16384));
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; i += 16) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int k = u & 3;
                const uint32_t d = tmem + (nacc == 1 ? 0 : nacc == 2 ? (u & 1) * 64 : (u & 3) * 64);
                const uint32_t accum = (i | (u >= nacc)) != 0;
                if (ts)
                    umma_f16_ts(d, tmem + 256 + k * 8, db + 2 * k, idesc, accum);
                else
                    umma_f16(d, da + 2 * k, db + 2 * k, idesc, accum);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static double avg(const long long* c, int n) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += static_cast<double>(c[i]);
    return s / n;
}

int main() {
    const int grid = 148;
    long long* d_cyc;
    float* d_sink;
    cudaMalloc(&d_cyc, grid * sizeof(long long));
    cudaMalloc(&d_sink, 4);
    long long h[148];
    const int iters = 256;
    const char* names[4] = {"tcgen05.ld 32x32b.x32", "tcgen05.st 32x32b.x32", "ex2.approx", "ld+ex2+pack+st (softmax-like)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) k_tmem<<<grid, warps * 32>>>(mode, iters, d_cyc, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("%s warps=%d failed: %s\n", names[mode], warps, cudaGetErrorString(e));
                return 1;
            }
            cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
            const double c = avg(h, grid);
            if (mode == 2)
                printf("%-32s warps=%2d  %9.0f cyc  %.2f ex2/cyc/SM\n", names[mode], warps, c,
                       double(iters) * 4 * 8 * 32 * warps / c);
            else if (mode == 3)
                printf("%-32s warps=%2d  %9.0f cyc  %.2f elements/cyc/SM\n", names[mode], warps, c,
                       double(iters) * 32 * 32 * warps / c);
            else
                printf("%-32s warps=%2d  %9.0f cyc  %.1f B/cyc/SM\n", names[mode], warps, c,
                       double(iters) * 4096 * warps / c);
        }
    cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int ts = 0; ts < 2; ++ts)
        for (int nacc : {1, 2, 4})
            for (int n : {64, 128, 256}) {
                const int b_mn = 0;
                if (nacc > 1 && n != 64) continue;
                const int n_mma = 512;
                for (int rep = 0; rep < 2; ++rep) k_mma<<<grid, 128, 64 * 1024>>>(n, ts, b_mn, n_mma, nacc, d_cyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("mma n=%d ts=%d b_mn=%d failed: %s\n", n, ts, b_mn, cudaGetErrorString(e));
                    return 1;
                }
                cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
                const double c = avg(h, grid);
                printf("tcgen05.mma M=128 N=%3d K=16 A=%s accumulators=%d: %7.1f cyc/mma  %.0f FLOP/cyc/SM\n", n,
                       ts ? "tmem" : "smem", nacc, c / n_mma, 2.0 * 128 * n * 16 * n_mma / c);
            }
    return 0;
}
