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
                                                                    int rows) {
    constexpr int D = NV * 128;
    const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
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

// dx_io += LN_bwd(dy) ; dx_bf16 = bf16(dx_io).  All three input streams are requested before the first reduction.
template <int NV>
__global__ void __launch_bounds__(kRowThreads) layernorm_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ gamma,
                                                                    float* __restrict__ dx_io, bf16* __restrict__ dx_bf16,
                                                                    int rows) {
    constexpr int D = NV * 128;
    const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(row) * D);
    float4* dxr = reinterpret_cast<float4*>(dx_io + static_cast<size_t>(row) * D);
    float4 v[NV], g[NV], acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = xr[i * 32 + lane];
        g[i] = bf16x4_to_float4(dyr[i * 32 + lane]);
        acc[i] = dxr[i * 32 + lane];
    }
    ln_bwd_row<NV>(v, g, gamma, lane);
    uint2* dbr = reinterpret_cast<uint2*>(dx_bf16 + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 d = make_float4(acc[i].x + g[i].x, acc[i].y + g[i].y, acc[i].z + g[i].z, acc[i].w + g[i].w);
        dxr[i * 32 + lane] = d;
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
template <int NV>
__global__ void __launch_bounds__(kRowThreads) embed_bwd_kernel(const float* __restrict__ dx0, const float* __restrict__ v,
                                                                const float* __restrict__ gamma, bf16* __restrict__ d_patch,
                                                                int n, int T) {
    constexpr int D = NV * 128;
    const int prow = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);  // patch row
    const int lane = threadIdx.x & 31;
    if (prow >= n * (T - 1)) return;
    const int nn = prow / (T - 1), t = prow % (T - 1) + 1;
    const size_t row = static_cast<size_t>(nn) * T + t;
    const float4* vr = reinterpret_cast<const float4*>(v + row * D);
    const float4* dr = reinterpret_cast<const float4*>(dx0 + row * D);
    float4 xv[NV], g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        xv[i] = vr[i * 32 + lane];
        g[i] = dr[i * 32 + lane];
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
        case 6: KERNEL<6><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                             \
        case 8: KERNEL<8><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                             \
        case 10: KERNEL<10><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                           \
        case 12: KERNEL<12><<<GRID, kRowThreads, 0, STREAM>>>(__VA_ARGS__); break;                           \
        default: return ::pcg::set_error(-1, "width %d is not one of 128*{1,2,4,6,8,10,12}", (D));           \
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

__global__ void __launch_bounds__(kHeadThreads)
head_loss_kernel(const float* __restrict__ x, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                 const float* __restrict__ proj, const float* __restrict__ targets, const float* __restrict__ tweights,
                 int T, int D, int E, int M, float scale, int normalize, float* __restrict__ loss_sum,
                 float* __restrict__ enc_out, const float* __restrict__ d_enc, float* __restrict__ dx,
                 bf16* __restrict__ dx_bf16) {
    extern __shared__ float sm[];
    float* xhat = sm;        // [D]
    float* ybuf = sm + D;    // [D]   ln_post output, later d(ln_post output)
    float* ebuf = sm + 2 * D;       // [E] z, then e
    float* gbuf = sm + 2 * D + E;   // [E] de, then dz
    __shared__ float red[kHeadThreads / 32];

    const int n = blockIdx.x;
    const int tid = threadIdx.x;
    const float* xr = x + static_cast<size_t>(n) * T * D;  // CLS row

    // ln_post
    float s = 0.f;
    for (int d = tid; d < D; d += kHeadThreads) s += xr[d];
    const float mean = block_sum(s, red) / D;
    float q = 0.f;
    for (int d = tid; d < D; d += kHeadThreads) {
        const float c = xr[d] - mean;
        q += c * c;
    }
    const float rstd = 1.0f / sqrtf(block_sum(q, red) / D + kLnEps);
    for (int d = tid; d < D; d += kHeadThreads) {
        const float xh = (xr[d] - mean) * rstd;
        xhat[d] = xh;
        ybuf[d] = xh * ln_g[d] + ln_b[d];
    }
    __syncthreads();
    // z = y @ proj   (proj [D,E] row-major: threads sweep e, coalesced)
    for (int e = tid; e < E; e += kHeadThreads) {
        float acc = 0.f;
#pragma unroll 4
        for (int d = 0; d < D; ++d) acc = fmaf(ybuf[d], __ldg(proj + static_cast<size_t>(d) * E + e), acc);
        ebuf[e] = acc;
    }
    __syncthreads();
    float zz = 0.f;
    for (int e = tid; e < E; e += kHeadThreads) zz += ebuf[e] * ebuf[e];
    const float znorm = sqrtf(block_sum(zz, red));
    const float inv_norm = normalize ? 1.0f / fmaxf(znorm, 1e-12f) : 1.0f;
    for (int e = tid; e < E; e += kHeadThreads) {
        const float ev = ebuf[e] * inv_norm;
        ebuf[e] = ev;
        gbuf[e] = 0.f;
        if (enc_out != nullptr) enc_out[static_cast<size_t>(n) * E + e] = ev;
    }
    __syncthreads();
    if ((targets == nullptr || M <= 0) && d_enc == nullptr) return;
    if (d_enc != nullptr) {
        for (int e = tid; e < E; e += kHeadThreads) gbuf[e] = d_enc[static_cast<size_t>(n) * E + e];
        M = 0;  // upstream gradient replaces the loss gradient
    }

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
        const float half = fminf(0.5f * r, 1.0f);
        const float theta = asinf(half);
        const float w = tweights[m];
        loss_local += w * 2.0f * theta * theta;
        // d/de = 2*theta / sqrt(1 - r^2/4) * (e - t) / r ; subgradient 0 at r == 0 (torch.norm backward)
        const float denom = sqrtf(fmaxf(1.0f - half * half, 1e-12f)) * r;
        const float coef = (r > 0.f) ? w * 2.0f * theta / denom : 0.f;
        for (int e = tid; e < E; e += kHeadThreads) gbuf[e] += coef * (ebuf[e] - tm[e]);
    }
    if (tid == 0 && loss_sum != nullptr && M > 0) atomicAdd(loss_sum, loss_local * scale);
    if (dx == nullptr) return;
    __syncthreads();

    // through F.normalize: dz = (de - e (e . de)) / |z|
    if (normalize) {
        float dot = 0.f;
        for (int e = tid; e < E; e += kHeadThreads) dot += ebuf[e] * gbuf[e];
        dot = block_sum(dot, red);
        for (int e = tid; e < E; e += kHeadThreads) gbuf[e] = (gbuf[e] - ebuf[e] * dot) * inv_norm * scale;
    } else {
        for (int e = tid; e < E; e += kHeadThreads) gbuf[e] *= scale;
    }
    __syncthreads();
    // dy = proj @ dz   (one warp per d: lanes sweep e, coalesced)
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int d = warp; d < D; d += kHeadThreads / 32) {
            const float* pr = proj + static_cast<size_t>(d) * E;
            float acc = 0.f;
            for (int e = lane; e < E; e += 32) acc = fmaf(__ldg(pr + e), gbuf[e], acc);
            acc = warp_sum(acc);
            if (lane == 0) ybuf[d] = acc * ln_g[d];  // g = dy * gamma
        }
    }
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int d = tid; d < D; d += kHeadThreads) {
        s1 += ybuf[d];
        s2 += ybuf[d] * xhat[d];
    }
    const float c1 = block_sum(s1, red) / D;
    const float c2 = block_sum(s2, red) / D;
    float* dxr = dx + static_cast<size_t>(n) * T * D;
    bf16* dbr = dx_bf16 ? dx_bf16 + static_cast<size_t>(n) * T * D : nullptr;
    for (int d = tid; d < D; d += kHeadThreads) {
        const float gval = rstd * (ybuf[d] - c1 - xhat[d] * c2);
        dxr[d] = gval;
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

extern "C" int pcg_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, int rows, int D,
                                 void* stream) {
    PCG_CHECK_ARG(x && gamma && beta && y_bf16 && rows > 0, "pcg_layernorm_fwd: bad arguments");
    if (int rc = check_d("pcg_layernorm_fwd", D)) return rc;
    ProfileScope prof(PCG_PROF_LAYERNORM, 6.0 * rows * D, static_cast<cudaStream_t>(stream));
    PCG_DISPATCH_NV(D, layernorm_fwd_kernel, ceil_div(rows, kRowsPerBlock), static_cast<cudaStream_t>(stream), x, gamma, beta,
                    static_cast<bf16*>(y_bf16), rows);
    PCG_LAUNCH_CHECK("layernorm_fwd_kernel");
    return 0;
}

extern "C" int pcg_layernorm_bwd(const void* dy_bf16, const float* x, const float* gamma, float* dx_io, void* dx_bf16,
                                 int rows, int D, void* stream) {
    PCG_CHECK_ARG(dy_bf16 && x && gamma && dx_io && dx_bf16 && rows > 0, "pcg_layernorm_bwd: bad arguments");
    if (int rc = check_d("pcg_layernorm_bwd", D)) return rc;
    ProfileScope prof(PCG_PROF_LAYERNORM, 16.0 * rows * D, static_cast<cudaStream_t>(stream));
    PCG_DISPATCH_NV(D, layernorm_bwd_kernel, ceil_div(rows, kRowsPerBlock), static_cast<cudaStream_t>(stream),
                    static_cast<const bf16*>(dy_bf16), x, gamma, dx_io, static_cast<bf16*>(dx_bf16), rows);
    PCG_LAUNCH_CHECK("layernorm_bwd_kernel");
    return 0;
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

extern "C" int pcg_embed_bwd(const float* dx0, const float* v, const float* gamma, void* d_patch_bf16, int n, int T,
                             int D, void* stream) {
    PCG_CHECK_ARG(dx0 && v && gamma && d_patch_bf16 && n > 0 && T > 1, "pcg_embed_bwd: bad arguments");
    if (int rc = check_d("pcg_embed_bwd", D)) return rc;
    ProfileScope prof(PCG_PROF_EMBED, 10.0 * n * T * D, static_cast<cudaStream_t>(stream));
    PCG_DISPATCH_NV(D, embed_bwd_kernel, ceil_div(n * (T - 1), kRowsPerBlock), static_cast<cudaStream_t>(stream), dx0, v,
                    gamma, static_cast<bf16*>(d_patch_bf16), n, T);
    PCG_LAUNCH_CHECK("embed_bwd_kernel");
    return 0;
}

extern "C" int pcg_head_loss(const float* x, const float* ln_g, const float* ln_b, const float* proj,
                             const float* targets, const float* tweights, int n, int T, int D, int E, int M, float scale,
                             int normalize, float* loss_sum, float* enc_out, const float* d_enc, float* dx,
                             void* dx_bf16, void* stream) {
    PCG_CHECK_ARG(x && ln_g && ln_b && proj && n > 0 && T > 0 && D > 0 && E > 0, "pcg_head_loss: bad arguments");
    PCG_CHECK_ARG(M == 0 || (targets && tweights), "pcg_head_loss: targets/tweights missing for M=%d", M);
    PCG_CHECK_ARG(dx == nullptr || M > 0 || d_enc, "pcg_head_loss: a gradient needs targets or d_enc");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(2 * D + 2 * E) * sizeof(float);
    PCG_CHECK_ARG(smem <= 48 * 1024, "pcg_head_loss: D=%d E=%d exceed the shared-memory budget", D, E);
    ProfileScope prof(PCG_PROF_HEAD, (dx ? 6.0 : 0.0) * n * T * D + 8.0 * n * D * E, s);
    if (dx != nullptr) {
        PCG_CUDA(cudaMemsetAsync(dx, 0, static_cast<size_t>(n) * T * D * sizeof(float), s));
        if (dx_bf16 != nullptr) PCG_CUDA(cudaMemsetAsync(dx_bf16, 0, static_cast<size_t>(n) * T * D * 2, s));
    }
    head_loss_kernel<<<n, kHeadThreads, smem, s>>>(x, ln_g, ln_b, proj, M > 0 ? targets : nullptr, tweights, T, D, E, M,
                                                   scale, normalize, loss_sum, enc_out, d_enc, dx, static_cast<bf16*>(dx_bf16));
    PCG_LAUNCH_CHECK("head_loss_kernel");
    return 0;
}
