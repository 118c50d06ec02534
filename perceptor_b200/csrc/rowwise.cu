// Row-wise HBM-bound kernels of the ViT: LayerNorm forward/backward on the fp32 residual stream (bf16 GEMM operand
// out), class-token + positional-embedding + ln_pre assembly, and the fused head
// (ln_post(CLS) @ proj -> L2 normalise -> spherical-distance loss, forward and analytic gradient in one launch).
// One warp per row, float4 loads issued up front (all of a row's traffic is in flight before the first reduction),
// warp-shuffle reductions, fp32 statistics (recomputed in backward, not stored).  Templated on D/128 so that the
// per-lane row slice lives in exactly-sized register arrays.
//
// Replaces LayerNorm (perceptor/models/ruclip/model.py:11-17), the cat/pos-emb/ln_pre sequence (:109-120),
// ln_post + proj (:126-129), F.normalize (perceptor/models/open_clip.py:120-121) and CLIP.forward's distance
// (perceptor/losses/clip/clip.py:89-99).
#include "pcg_common.cuh"
#include "pcg_ptx.cuh"

namespace pcg {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kMaxV = 12;  // float4 per lane: supports D <= 1536, D % 128 == 0
constexpr float kLnEps = 1e-5f;
constexpr int kRowThreads = 128;  // 4 rows per block
constexpr int kRowsPerBlock = kRowThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

struct RowStats {
    float mean, rstd;
};

// x[] holds this lane's float4s; returns mean / rstd of the whole row (biased variance, eps 1e-5)
template <int NV>
__device__ __forceinline__ RowStats row_stats(const float4 (&x)[NV]) {
    constexpr float inv_d = 1.0f / (NV * 128);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float var = warp_sum(q) * inv_d;
    return {mean, 1.0f / sqrtf(var + kLnEps)};
}

__device__ __forceinline__ uint2 pack4(float a, float b, float c, float d) {
    return make_uint2(pack_bf16(a, b), pack_bf16(c, d));
}

// --------------------------------------------------------------------------------------------------------
// LayerNorm forward: y(bf16) = (x - mean) * rstd * gamma + beta
// --------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kRowThreads) layernorm_fwd_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, bf16* __restrict__ y,
                                                                    int rows, int row_step) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    constexpr int D = NV * 128;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5) >= rows) return;
    // row_step > 1: every row_step-th row of x and y (the class-token rows of the last block)
#ifndef PCG_EXP_LN_FORWARD_ORDER
    // Last rows first: the producing GEMM wrote its [M, D] output top to bottom and M * D * 4 bytes exceed the L2, so
    // the rows still resident are the last ones; walking the same direction would miss on every row (LRU), and this
    // kernel's own bf16 output then ends with the rows the next GEMM reads first.  Measured: 4.64 -> 4.53 ms of
    // LayerNorm per ViT-L/14 step and 0.4 ms off the step.
    const size_t row = static_cast<size_t>(rows - 1 - static_cast<int>(blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5))) * row_step;
#else
    const size_t row = static_cast<size_t>(blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5)) * row_step;
#endif
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
    const RowStats st = row_stats<NV>(v);
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
        yr[i * 32 + lane] = pack4((v[i].x - st.mean) * st.rstd * g.x + b.x, (v[i].y - st.mean) * st.rstd * g.y + b.y,
                                  (v[i].z - st.mean) * st.rstd * g.z + b.z, (v[i].w - st.mean) * st.rstd * g.w + b.w);
    }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
template <int NV>
__device__ __forceinline__ void ln_bwd_row(const float4 (&x)[NV], float4 (&g)[NV], const float* gamma, int lane) {
    constexpr float inv_d = 1.0f / (NV * 128);
    const RowStats st = row_stats<NV>(x);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
        g[i].x *= gm.x; g[i].y *= gm.y; g[i].z *= gm.z; g[i].w *= gm.w;
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += g[i].x * (x[i].x - st.mean) + g[i].y * (x[i].y - st.mean) + g[i].z * (x[i].z - st.mean) +
              g[i].w * (x[i].w - st.mean);
    }
    const float c1 = warp_sum(s1) * inv_d;
    const float c2 = warp_sum(s2) * st.rstd * inv_d;  // mean(g * xhat)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        g[i].x = st.rstd * (g[i].x - c1 - (x[i].x - st.mean) * st.rstd * c2);
        g[i].y = st.rstd * (g[i].y - c1 - (x[i].y - st.mean) * st.rstd * c2);
        g[i].z = st.rstd * (g[i].z - c1 - (x[i].z - st.mean) * st.rstd * c2);
        g[i].w = st.rstd * (g[i].w - c1 - (x[i].w - st.mean) * st.rstd * c2);
    }
}

__device__ __forceinline__ float4 bf16x4_to_float4(uint2 u) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

// dx += LN_bwd(dy).  kF32: the residual gradient lives in fp32 (dx_io, 16 B/element) with a bf16 copy for the next
// dgrad GEMM; !kF32: the bf16 tensor IS the residual gradient, accumulated in place (10 B/element) -- the production
// backward: 48 bf16 roundings cost ~1e-4 of gradient cosine (tolerance 1e-3) and 37 % of this kernel's traffic.
// All input streams are requested before the first reduction.
template <int NV, bool kF32>
__global__ void __launch_bounds__(kRowThreads) layernorm_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ gamma,
                                                                    float* __restrict__ dx_io, bf16* __restrict__ dx_bf16,
                                                                    int rows, int row_step) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    constexpr int D = NV * 128;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5) >= rows) return;
#ifndef PCG_EXP_LN_FORWARD_ORDER  // last rows first, see layernorm_fwd_kernel
    const size_t row = static_cast<size_t>(rows - 1 - static_cast<int>(blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5))) * row_step;
#else
    const size_t row = static_cast<size_t>(blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5)) * row_step;
#endif
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(row) * D);
    float4* dxr = reinterpret_cast<float4*>(dx_io + static_cast<size_t>(row) * D);
    uint2* dbr = reinterpret_cast<uint2*>(dx_bf16 + static_cast<size_t>(row) * D);
    float4 v[NV], g[NV], acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = xr[i * 32 + lane];
        g[i] = bf16x4_to_float4(dyr[i * 32 + lane]);
        if constexpr (kF32) acc[i] = dxr[i * 32 + lane];
        else acc[i] = bf16x4_to_float4(dbr[i * 32 + lane]);
    }
    ln_bwd_row<NV>(v, g, gamma, lane);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 d = make_float4(acc[i].x + g[i].x, acc[i].y + g[i].y, acc[i].z + g[i].z, acc[i].w + g[i].w);
        if constexpr (kF32) dxr[i * 32 + lane] = d;
        dbr[i * 32 + lane] = pack4(d.x, d.y, d.z, d.w);
    }
}

// --------------------------------------------------------------------------------------------------------
// embed: v = (t == 0 ? cls : patch_out[n*g*g + t-1]) + pos[t];  x0 = ln_pre(v)
// --------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kRowThreads) embed_fwd_kernel(const float* __restrict__ patch_out,
                                                                const float* __restrict__ cls, const float* __restrict__ pos,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ vout, float* __restrict__ x0, int n, int T) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    constexpr int D = NV * 128;
    const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n * T) return;
    const int nn = row / T, t = row % T;
    const float4* src = (t == 0) ? reinterpret_cast<const float4*>(cls)
                                 : reinterpret_cast<const float4*>(patch_out + (static_cast<size_t>(nn) * (T - 1) + (t - 1)) * D);
    const float4* pr = reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * D);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 a = src[i * 32 + lane];
        const float4 b = __ldg(pr + i * 32 + lane);
        v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    const RowStats st = row_stats<NV>(v);
    float4* vr = reinterpret_cast<float4*>(vout + static_cast<size_t>(row) * D);
    float4* xr = reinterpret_cast<float4*>(x0 + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
        vr[i * 32 + lane] = v[i];
        xr[i * 32 + lane] = make_float4((v[i].x - st.mean) * st.rstd * g.x + b.x, (v[i].y - st.mean) * st.rstd * g.y + b.y,
                                        (v[i].z - st.mean) * st.rstd * g.z + b.z, (v[i].w - st.mean) * st.rstd * g.w + b.w);
    }
}

// d_patch[n*g*g + t-1] = bf16( ln_pre'(v)^T dx0[n*T + t] ), t >= 1
template <int NV, bool kF32>
__global__ void __launch_bounds__(kRowThreads) embed_bwd_kernel(const void* __restrict__ dx0, const float* __restrict__ v,
                                                                const float* __restrict__ gamma, bf16* __restrict__ d_patch,
                                                                int n, int T) {
    constexpr int D = NV * 128;
    const int prow = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);  // patch row
    const int lane = threadIdx.x & 31;
    if (prow >= n * (T - 1)) return;
    const int nn = prow / (T - 1), t = prow % (T - 1) + 1;
    const size_t row = static_cast<size_t>(nn) * T + t;
    const float4* vr = reinterpret_cast<const float4*>(v + row * D);
    float4 xv[NV], g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        xv[i] = vr[i * 32 + lane];
        if constexpr (kF32) g[i] = reinterpret_cast<const float4*>(static_cast<const float*>(dx0) + row * D)[i * 32 + lane];
        else g[i] = bf16x4_to_float4(reinterpret_cast<const uint2*>(static_cast<const bf16*>(dx0) + row * D)[i * 32 + lane]);
    }
    ln_bwd_row<NV>(xv, g, gamma, lane);
    uint2* out = reinterpret_cast<uint2*>(d_patch + static_cast<size_t>(prow) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) out[i * 32 + lane] = pack4(g[i].x, g[i].y, g[i].z, g[i].w);
}

// launch KERNEL<NV>(args...) for the supported widths D = 128 * NV
#define PCG_DISPATCH_NV(D, KERNEL, GRID, STREAM, ...)                                                        \
    switch ((D) >> 7) {                                                                                      \
        case 1: KERNEL<1><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                             \
        case 2: KERNEL<2><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                             \
        case 4: KERNEL<4><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                             \
        case 5: KERNEL<5><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                             \
        case 6: KERNEL<6><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                             \
        case 8: KERNEL<8><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                             \
        case 10: KERNEL<10><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                           \
        case 11: KERNEL<11><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                           \
        case 12: KERNEL<12><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                           \
        default: return ::pcg::set_error(-1, "width %d is not one of 128*{1,2,4,5,6,8,10,11,12}", (D));           \
    }

// the same for kernels templated on <NV, bool>
#define PCG_DISPATCH_NV2(D, KERNEL, FLAG, GRID, STREAM, ...)                                                 \
    switch ((D) >> 7) {                                                                                      \
        case 1: KERNEL<1, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                       \
        case 2: KERNEL<2, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                       \
        case 4: KERNEL<4, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                       \
        case 5: KERNEL<5, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                       \
        case 6: KERNEL<6, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                       \
        case 8: KERNEL<8, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                       \
        case 10: KERNEL<10, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                     \
        case 11: KERNEL<11, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                     \
        case 12: KERNEL<12, FLAG><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                     \
        default: return ::pcg::set_error(-1, "width %d is not one of 128*{1,2,4,5,6,8,10,11,12}", (D));           \
    }

// --------------------------------------------------------------------------------------------------------
// head: one CTA per cutout.
// --------------------------------------------------------------------------------------------------------
constexpr int kHeadThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from the previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadThreads / 32; ++i) t += red[i];
    return t;
}

// The head is five small launches batched over cutouts, so that the projection matrix (3 MB for ViT-L/14) is read
// once per 16-cutout tile instead of twice per cutout:
//   head_ln_kernel      y  = ln_post(x[CLS])                       one CTA per cutout
//   head_gemm_kernel    z  = y @ proj                              16 x 64 output tiles
//   head_dist_kernel    e = z / |z|, loss, dz = d(loss)/dz         one CTA per cutout
//   head_gemm_kernel    dy = dz @ proj^T
//   head_ln_bwd_kernel  dx[CLS] = ln_post'(dy)                     one CTA per cutout
__global__ void __launch_bounds__(kHeadThreads)
head_ln_kernel(const float* __restrict__ x, const float* __restrict__ ln_g, const float* __restrict__ ln_b, int T, int D,
               float* __restrict__ y) {
    __shared__ float red[kHeadThreads / 32];
    const int n = blockIdx.x, tid = threadIdx.x;
    const float* xr = x + static_cast<size_t>(n) * T * D;  // CLS row
    float s = 0.f;
    for (int d = tid; d < D; d += kHeadThreads) s += xr[d];
    const float mean = block_sum(s, red) / D;
    float q = 0.f;
    for (int d = tid; d < D; d += kHeadThreads) {
        const float c = xr[d] - mean;
        q += c * c;
    }
    const float rstd = 1.0f / sqrtf(block_sum(q, red) / D + kLnEps);
    for (int d = tid; d < D; d += kHeadThreads) y[static_cast<size_t>(n) * D + d] = (xr[d] - mean) * rstd * ln_g[d] + ln_b[d];
}

// C[n, N] = A[n, K] @ B with B = Bm[K, N] (kTransB = false) or B = Bm[N, K]^T (kTransB = true), fp32 on the CUDA
// cores: 16 x 64 tile per CTA, 32-deep K chunks through shared memory, thread = 1 row x 4 columns.  The problem is
// tiny (128 x 768 x 1024 for ViT-L/14) and latency bound, so K is split kHgParts ways over blockIdx.z to put four
// times as many CTAs to work; part z writes its partial sums to C + z * n * N and the consumer kernel adds them up
// (a fixed order: deterministic, unlike atomics).
constexpr int kHgRows = 16, kHgCols = 64, kHgK = 32, kHgParts = 4;
template <bool kTransB>
__global__ void __launch_bounds__(kHeadThreads)
head_gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, int n, int N, int K) {
    __shared__ float As[kHgRows][kHgK + 1];
    __shared__ float Bs[kHgK][kHgCols + 4];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.y * kHgRows, c0 = blockIdx.x * kHgCols;
    const int row = tid >> 4, col = (tid & 15) * 4;
    const int k_per = ((K + kHgParts - 1) / kHgParts + kHgK - 1) / kHgK * kHgK;
    const int k_begin = blockIdx.z * k_per, k_end = min(K, k_begin + k_per);
    C += static_cast<size_t>(blockIdx.z) * n * N;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = k_begin; k0 < k_end; k0 += kHgK) {
        for (int i = tid; i < kHgRows * kHgK; i += kHeadThreads) {
            const int r = i / kHgK, k = i - r * kHgK;
            As[r][k] = (n0 + r < n && k0 + k < k_end) ? A[static_cast<size_t>(n0 + r) * K + k0 + k] : 0.f;
        }
        for (int i = tid; i < kHgK * kHgCols; i += kHeadThreads) {
            if (kTransB) {  // Bm rows index the output column, contiguous along k
                const int c = i / kHgK, k = i - c * kHgK;
                Bs[k][c] = (c0 + c < N && k0 + k < k_end) ? __ldg(Bm + static_cast<size_t>(c0 + c) * K + k0 + k) : 0.f;
            } else {
                const int k = i / kHgCols, c = i - k * kHgCols;
                Bs[k][c] = (c0 + c < N && k0 + k < k_end) ? __ldg(Bm + static_cast<size_t>(k0 + k) * N + c0 + c) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kHgK; ++k) {
            const float a = As[row][k];
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][col]);
            acc[0] = fmaf(a, b.x, acc[0]);
            acc[1] = fmaf(a, b.y, acc[1]);
            acc[2] = fmaf(a, b.z, acc[2]);
            acc[3] = fmaf(a, b.w, acc[3]);
        }
        __syncthreads();
    }
    if (n0 + row < n) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c0 + col + j < N) C[static_cast<size_t>(n0 + row) * N + c0 + col + j] = acc[j];
    }
}

// row `n` of a split-K result: add parts 1 .. kHgParts-1 into part 0 (each CTA owns one row)
__device__ __forceinline__ void combine_parts(float* base, int n_rows, int width, int n) {
    float* r0 = base + static_cast<size_t>(n) * width;
    for (int e = threadIdx.x; e < width; e += kHeadThreads) {
        float v = r0[e];
#pragma unroll
        for (int q = 1; q < kHgParts; ++q) v += base[(static_cast<size_t>(q) * n_rows + n) * width + e];
        r0[e] = v;
    }
    __syncthreads();
}

// e = z / |z| (or z), spherical-distance loss against every target and dz = d(loss * scale)/dz (or the pull-back of
// an upstream gradient d_enc through the normalisation)
__global__ void __launch_bounds__(kHeadThreads)
head_dist_kernel(float* __restrict__ z, const float* __restrict__ targets, const float* __restrict__ tweights,
                 int E, int M, float scale, int normalize, float* __restrict__ loss_part, float* __restrict__ enc_out,
                 const float* __restrict__ d_enc, float* __restrict__ dz) {
    extern __shared__ float sm[];
    float* ebuf = sm;      // [E]
    float* gbuf = sm + E;  // [E] d/de
    __shared__ float red[kHeadThreads / 32];
    const int n = blockIdx.x, tid = threadIdx.x;
    combine_parts(z, gridDim.x, E, n);  // z arrives as kHgParts split-K partials
    const float* zr = z + static_cast<size_t>(n) * E;
    float zz = 0.f;
    for (int e = tid; e < E; e += kHeadThreads) zz += zr[e] * zr[e];
    const float znorm = sqrtf(block_sum(zz, red));
    const float inv_norm = normalize ? 1.0f / fmaxf(znorm, 1e-12f) : 1.0f;
    for (int e = tid; e < E; e += kHeadThreads) {
        const float ev = zr[e] * inv_norm;
        ebuf[e] = ev;
        gbuf[e] = (d_enc != nullptr) ? d_enc[static_cast<size_t>(n) * E + e] : 0.f;
        if (enc_out != nullptr) enc_out[static_cast<size_t>(n) * E + e] = ev;
    }
    __syncthreads();
    if (d_enc != nullptr) M = 0;  // the upstream gradient replaces the loss gradient
    // spherical distance to every target: d = 2 * asin(r/2)^2, r = |e - t|
    float loss_local = 0.f;
    for (int m = 0; m < M; ++m) {
        const float* tm = targets + static_cast<size_t>(m) * E;
        float rr = 0.f;
        for (int e = tid; e < E; e += kHeadThreads) {
            const float df = ebuf[e] - tm[e];
            rr += df * df;
        }
        rr = block_sum(rr, red);
        const float r = sqrtf(rr);
        // r / 2 > 1 cannot happen between unit vectors; it does when a caller stores un-normalised targets (OpenCLIP
        // keeps them as given, perceptor/losses/open_clip.py:58-85).  The reference then returns NaN (asin of > 1)
        // and so does this: a finite loss with an exploding gradient would hide the mistake.  Rounding slop just
        // above 1 (antipodal unit vectors) is treated as exactly 1.
        const float half_raw = 0.5f * r;
        const float half = !(half_raw <= 1.000001f) ? __int_as_float(0x7fc00000) : fminf(half_raw, 1.0f);  // NaN in, NaN out
        const float theta = asinf(half);
        const float w = tweights[m];
        loss_local += w * 2.0f * theta * theta;
        // d/de = 2*theta / sqrt(1 - r^2/4) * (e - t) / r ; subgradient 0 at r == 0 (torch.norm backward) and at the
        // antipode r == 2, where the derivative of asin is unbounded
        const float one_minus = 1.0f - half * half;
        const float coef = (r == 0.f || one_minus <= 1e-12f) ? 0.f : w * 2.0f * theta / (sqrtf(one_minus) * r);
        for (int e = tid; e < E; e += kHeadThreads) gbuf[e] += coef * (ebuf[e] - tm[e]);
    }
    // per-cutout partial; head_loss_sum_kernel adds them in a fixed order (bit-reproducible, unlike atomics)
    if (tid == 0 && loss_part != nullptr) loss_part[n] = loss_local * scale;
    if (dz == nullptr) return;
    __syncthreads();
    // through F.normalize: dz = (de - e (e . de)) / |z|
    float dot = 0.f;
    if (normalize) {
        for (int e = tid; e < E; e += kHeadThreads) dot += ebuf[e] * gbuf[e];
        dot = block_sum(dot, red);
    }
    for (int e = tid; e < E; e += kHeadThreads)
        dz[static_cast<size_t>(n) * E + e] = normalize ? (gbuf[e] - ebuf[e] * dot) * inv_norm * scale : gbuf[e] * scale;
}

// *loss_sum += sum_n loss_part[n], always in the same order: thread t adds elements t, t + 256, ... sequentially, then
// the block tree-reduces
__global__ void __launch_bounds__(kHeadThreads) head_loss_sum_kernel(const float* __restrict__ loss_part, int n,
                                                                     float* __restrict__ loss_sum) {
    __shared__ float red[kHeadThreads / 32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += kHeadThreads) s += loss_part[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *loss_sum += s;
}

// dx[CLS row] = ln_post'(dy): statistics recomputed from x (one 4 KB row)
__global__ void __launch_bounds__(kHeadThreads)
head_ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ ln_g, float* __restrict__ dy, int T, int D,
                   float* __restrict__ dx, bf16* __restrict__ dx_bf16) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    __shared__ float red[kHeadThreads / 32];
    const int n = blockIdx.x, tid = threadIdx.x;
    combine_parts(dy, gridDim.x, D, n);  // dy arrives as kHgParts split-K partials
    const float* xr = x + static_cast<size_t>(n) * T * D;
    const float* gr = dy + static_cast<size_t>(n) * D;
    float s = 0.f;
    for (int d = tid; d < D; d += kHeadThreads) s += xr[d];
    const float mean = block_sum(s, red) / D;
    float q = 0.f;
    for (int d = tid; d < D; d += kHeadThreads) {
        const float c = xr[d] - mean;
        q += c * c;
    }
    const float rstd = 1.0f / sqrtf(block_sum(q, red) / D + kLnEps);
    float s1 = 0.f, s2 = 0.f;
    for (int d = tid; d < D; d += kHeadThreads) {
        const float g = gr[d] * ln_g[d];
        s1 += g;
        s2 += g * (xr[d] - mean) * rstd;
    }
    const float c1 = block_sum(s1, red) / D;
    const float c2 = block_sum(s2, red) / D;
    float* dxr = dx ? dx + static_cast<size_t>(n) * T * D : nullptr;
    bf16* dbr = dx_bf16 ? dx_bf16 + static_cast<size_t>(n) * T * D : nullptr;
    for (int d = tid; d < D; d += kHeadThreads) {
        const float xh = (xr[d] - mean) * rstd;
        const float gval = rstd * (gr[d] * ln_g[d] - c1 - xh * c2);
        if (dxr) dxr[d] = gval;
        if (dbr) dbr[d] = __float2bfloat16(gval);
    }
}

}  // namespace
}  // namespace pcg

using namespace pcg;

static int check_d(const char* who, int D) {
    if (D <= 0 || D % 128 != 0 || D > 128 * kMaxV) return set_error(-1, "%s: width %d must be a multiple of 128, <= %d", who, D, 128 * kMaxV);
    return 0;
}

extern "C" int pcg_layernorm_fwd_rows(const float* x, const float* gamma, const float* beta, void* y_bf16, int rows, int D,
                                      int row_step, void* stream) {
    PCG_CHECK_ARG(x && gamma && beta && y_bf16 && rows > 0 && row_step > 0, "pcg_layernorm_fwd: bad arguments");
    if (int rc = check_d("pcg_layernorm_fwd", D)) return rc;
    ProfileScope prof(PCG_PROF_LAYERNORM, 6.0 * rows * D, static_cast<cudaStream_t>(stream));
    PCG_DISPATCH_NV(D, layernorm_fwd_kernel, ceil_div(rows, kRowsPerBlock), static_cast<cudaStream_t>(stream), x, gamma, beta,
                    static_cast<bf16*>(y_bf16), rows, row_step);
    PCG_LAUNCH_CHECK("layernorm_fwd_kernel");
    return 0;
}
extern "C" int pcg_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, int rows, int D,
                                 void* stream) {
    return pcg_layernorm_fwd_rows(x, gamma, beta, y_bf16, rows, D, 1, stream);
}

extern "C" int pcg_layernorm_bwd_rows(const void* dy_bf16, const float* x, const float* gamma, float* dx_io, void* dx_bf16,
                                      int rows, int D, int row_step, void* stream) {
    PCG_CHECK_ARG(dy_bf16 && x && gamma && dx_bf16 && rows > 0 && row_step > 0, "pcg_layernorm_bwd: bad arguments");
    if (int rc = check_d("pcg_layernorm_bwd", D)) return rc;
    ProfileScope prof(PCG_PROF_LAYERNORM, (dx_io ? 16.0 : 10.0) * rows * D, static_cast<cudaStream_t>(stream));
    if (dx_io != nullptr) {
        PCG_DISPATCH_NV2(D, layernorm_bwd_kernel, true, ceil_div(rows, kRowsPerBlock), static_cast<cudaStream_t>(stream),
                         static_cast<const bf16*>(dy_bf16), x, gamma, dx_io, static_cast<bf16*>(dx_bf16), rows, row_step);
    } else {
        PCG_DISPATCH_NV2(D, layernorm_bwd_kernel, false, ceil_div(rows, kRowsPerBlock), static_cast<cudaStream_t>(stream),
                         static_cast<const bf16*>(dy_bf16), x, gamma, dx_io, static_cast<bf16*>(dx_bf16), rows, row_step);
    }
    PCG_LAUNCH_CHECK("layernorm_bwd_kernel");
    return 0;
}
extern "C" int pcg_layernorm_bwd(const void* dy_bf16, const float* x, const float* gamma, float* dx_io, void* dx_bf16,
                                 int rows, int D, void* stream) {
    return pcg_layernorm_bwd_rows(dy_bf16, x, gamma, dx_io, dx_bf16, rows, D, 1, stream);
}

extern "C" int pcg_embed_fwd(const float* patch_out, const float* cls, const float* pos, const float* gamma,
                             const float* beta, float* v, float* x0, int n, int T, int D, void* stream) {
    PCG_CHECK_ARG(patch_out && cls && pos && gamma && beta && v && x0 && n > 0 && T > 1, "pcg_embed_fwd: bad arguments");
    if (int rc = check_d("pcg_embed_fwd", D)) return rc;
    ProfileScope prof(PCG_PROF_EMBED, 12.0 * n * T * D, static_cast<cudaStream_t>(stream));
    PCG_DISPATCH_NV(D, embed_fwd_kernel, ceil_div(n * T, kRowsPerBlock), static_cast<cudaStream_t>(stream), patch_out, cls,
                    pos, gamma, beta, v, x0, n, T);
    PCG_LAUNCH_CHECK("embed_fwd_kernel");
    return 0;
}

extern "C" int pcg_embed_bwd(const float* dx0, const void* dx0_bf16, const float* v, const float* gamma,
                             void* d_patch_bf16, int n, int T, int D, void* stream) {
    PCG_CHECK_ARG((dx0 != nullptr) != (dx0_bf16 != nullptr), "pcg_embed_bwd: pass exactly one of dx0 / dx0_bf16");
    PCG_CHECK_ARG(v && gamma && d_patch_bf16 && n > 0 && T > 1, "pcg_embed_bwd: bad arguments");
    if (int rc = check_d("pcg_embed_bwd", D)) return rc;
    ProfileScope prof(PCG_PROF_EMBED, (dx0 ? 10.0 : 8.0) * n * T * D, static_cast<cudaStream_t>(stream));
    if (dx0 != nullptr) {
        PCG_DISPATCH_NV2(D, embed_bwd_kernel, true, ceil_div(n * (T - 1), kRowsPerBlock), static_cast<cudaStream_t>(stream),
                         static_cast<const void*>(dx0), v, gamma, static_cast<bf16*>(d_patch_bf16), n, T);
    } else {
        PCG_DISPATCH_NV2(D, embed_bwd_kernel, false, ceil_div(n * (T - 1), kRowsPerBlock), static_cast<cudaStream_t>(stream),
                         dx0_bf16, v, gamma, static_cast<bf16*>(d_patch_bf16), n, T);
    }
    PCG_LAUNCH_CHECK("embed_bwd_kernel");
    return 0;
}

extern "C" size_t pcg_head_workspace_bytes(int n, int D, int E) {
    if (n <= 0 || D <= 0 || E <= 0) return 0;
    // y [n,D], dy [parts][n,D], z [parts][n,E], dz [n,E], per-cutout loss partials [n]
    return (static_cast<size_t>(n) * D * (1 + kHgParts) + static_cast<size_t>(n) * E * (1 + kHgParts) + n) * sizeof(float);
}

extern "C" int pcg_head_loss(const float* x, const float* ln_g, const float* ln_b, const float* proj,
                             const float* targets, const float* tweights, int n, int T, int D, int E, int M, float scale,
                             int normalize, float* loss_sum, float* enc_out, const float* d_enc, float* dx,
                             void* dx_bf16, float* workspace, void* stream) {
    PCG_CHECK_ARG(x && ln_g && ln_b && proj && workspace && n > 0 && T > 0 && D > 0 && E > 0,
                  "pcg_head_loss: bad arguments");
    PCG_CHECK_ARG(M == 0 || (targets && tweights), "pcg_head_loss: targets/tweights missing for M=%d", M);
    const bool want_dx = dx != nullptr || dx_bf16 != nullptr;
    PCG_CHECK_ARG(!want_dx || M > 0 || d_enc, "pcg_head_loss: a gradient needs targets or d_enc");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(2 * E) * sizeof(float);
    PCG_CHECK_ARG(smem <= 48 * 1024, "pcg_head_loss: E=%d exceeds the shared-memory budget", E);
    ProfileScope prof(PCG_PROF_HEAD, ((dx ? 4.0 : 0.0) + (dx_bf16 ? 2.0 : 0.0)) * n * T * D + (want_dx ? 4.0 : 2.0) * n * D * E, s);
    float* y = workspace;                                        // [n, D] ln_post output
    float* dy = y + static_cast<size_t>(n) * D;                  // [parts][n, D]
    float* z = dy + static_cast<size_t>(n) * D * kHgParts;       // [parts][n, E]
    float* dz = z + static_cast<size_t>(n) * E * kHgParts;       // [n, E]
    float* loss_part = dz + static_cast<size_t>(n) * E;          // [n]
    if (dx != nullptr) PCG_CUDA(cudaMemsetAsync(dx, 0, static_cast<size_t>(n) * T * D * sizeof(float), s));
    if (dx_bf16 != nullptr) PCG_CUDA(cudaMemsetAsync(dx_bf16, 0, static_cast<size_t>(n) * T * D * 2, s));
    head_ln_kernel<<<n, kHeadThreads, 0, s>>>(x, ln_g, ln_b, T, D, y);
    PCG_LAUNCH_CHECK("head_ln_kernel");
    head_gemm_kernel<false><<<dim3(ceil_div(E, kHgCols), ceil_div(n, kHgRows), kHgParts), kHeadThreads, 0, s>>>(y, proj, z, n, E, D);
    PCG_LAUNCH_CHECK("head_gemm_kernel");
    const bool has_loss = targets != nullptr && M > 0;
    if (!has_loss && d_enc == nullptr && enc_out == nullptr) return 0;
    head_dist_kernel<<<n, kHeadThreads, smem, s>>>(z, has_loss ? targets : nullptr, tweights, E, has_loss ? M : 0, scale,
                                                   normalize, (has_loss && loss_sum) ? loss_part : nullptr, enc_out, d_enc,
                                                   want_dx ? dz : nullptr);
    PCG_LAUNCH_CHECK("head_dist_kernel");
    if (has_loss && loss_sum != nullptr) {
        head_loss_sum_kernel<<<1, kHeadThreads, 0, s>>>(loss_part, n, loss_sum);
        PCG_LAUNCH_CHECK("head_loss_sum_kernel");
    }
    if (!want_dx) return 0;
    head_gemm_kernel<true><<<dim3(ceil_div(D, kHgCols), ceil_div(n, kHgRows), kHgParts), kHeadThreads, 0, s>>>(dz, proj, dy, n, D, E);
    PCG_LAUNCH_CHECK("head_gemm_kernel");
    head_ln_bwd_kernel<<<n, kHeadThreads, 0, s>>>(x, ln_g, dy, T, D, dx, static_cast<bf16*>(dx_bf16));
    PCG_LAUNCH_CHECK("head_ln_bwd_kernel");
    return 0;
}
