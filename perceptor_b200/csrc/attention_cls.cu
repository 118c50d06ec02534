// Class-token attention of the LAST transformer block: one query row (token 0) against all T keys.
//
// The guidance loss reads only ln_post(x[CLS]) (perceptor/models/ruclip/model.py:126-129), so in the final block
// everything after the K / V projections is needed for the class-token row alone: its attention output, the
// out-projection, ln_2 and the MLP.  The sequencer (api.cu) therefore runs the last block on n rows instead of n*T and
// calls these two kernels instead of the T x T attention; results are the same numbers the full block would put in
// the class-token row (the other rows of the last block's output are never read by anything).
//
// One CTA per (cutout, head), 4 warps.  A group of HS/8 lanes owns one key row (16 bytes per lane: every 128-byte
// line of K / V is requested whole), 32 / (HS/8) keys per warp step.  Both kernels are bound by the one pass over
// K and V (forward: 2 * T * HS * 2 bytes per CTA; backward reads the same and writes dQ / dK / dV rows of the head).
// HS is the column stride of a head in qkv: 64, or 128 for the padded head-dim-80/88 towers (pad columns are zero in
// q, k, v and dO, so they contribute nothing and receive zeros).
//
// Replaces, for the last block only, the nn.MultiheadAttention core of ruclip/model.py:43-49 and its autograd.
#include "pcg_common.cuh"
#include "pcg_ptx.cuh"

namespace pcg {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kClsThreads = 128;
constexpr int kClsWarps = kClsThreads / 32;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16), f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16), f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16), f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16), f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
    float s0 = a[0] * b[0], s1 = a[1] * b[1];
#pragma unroll
    for (int i = 2; i < 8; i += 2) s0 = fmaf(a[i], b[i], s0), s1 = fmaf(a[i + 1], b[i + 1], s1);
    return s0 + s1;
}
// sum over the LPR lanes of a key-row group (lanes differing in their low log2(LPR) bits)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int s = 1; s < LPR; s <<= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
// sum over the groups of a warp (lanes with the same column chunk)
template <int LPR>
__device__ __forceinline__ float across_groups_sum(float v) {
#pragma unroll
    for (int s = LPR; s < 32; s <<= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
__device__ __forceinline__ float warp_max_all(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
    return v;
}
__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// out[n*T + 0, h*HS : +HS] = softmax(q_0 K^T) V ; lse[n, h, 0] = log sum exp (natural log; q arrives pre-scaled)
template <int HS>
__global__ void __launch_bounds__(kClsThreads)
attn_cls_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int T, int heads) {
    grid_dep_launch();
    constexpr int LPR = HS / 8, KPW = 32 / LPR, KSTEP = KPW * kClsWarps;
    extern __shared__ float sm_cls[];
    float* sc = sm_cls;                    // [T] scores, then probabilities
    float* red = sm_cls + ((T + 3) & ~3);  // [kClsWarps][HS] partial outputs, then 2 * kClsWarps scalars
    float* red_s = red + kClsWarps * HS;
    const int h = blockIdx.x, n = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane % LPR, kslot = warp * KPW + lane / LPR;
    const size_t Da = static_cast<size_t>(heads) * HS, ld = 3 * Da;
    const bf16* base = qkv + static_cast<size_t>(n) * T * ld + static_cast<size_t>(h) * HS + sub * 8;
    float qf[8];
    unpack8(*reinterpret_cast<const uint4*>(base), qf);
    // the trip count is warp-uniform (the group sums shuffle across the whole warp): rows past T are predicated off
    for (int j0 = warp * KPW; j0 < T; j0 += KSTEP) {
        const int j = j0 + lane / LPR;
        const bool ok = j < T;
        float kf[8];
        unpack8(ok ? *reinterpret_cast<const uint4*>(base + j * ld + Da) : make_uint4(0u, 0u, 0u, 0u), kf);
        const float s = group_sum<LPR>(dot8(qf, kf));
        if (ok && sub == 0) sc[j] = s;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < T; j += kClsThreads) mx = fmaxf(mx, sc[j]);
    mx = warp_max_all(mx);
    if (lane == 0) red_s[warp] = mx;
    __syncthreads();
    mx = red_s[0];
#pragma unroll
    for (int w = 1; w < kClsWarps; ++w) mx = fmaxf(mx, red_s[w]);
    float sum = 0.f;
    for (int j = threadIdx.x; j < T; j += kClsThreads) {
        const float p = __expf(sc[j] - mx);
        sc[j] = p;
        sum += p;
    }
    sum = warp_sum_all(sum);
    if (lane == 0) red_s[kClsWarps + warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < kClsWarps; ++w) sum += red_s[kClsWarps + w];
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = kslot; j < T; j += KSTEP) {
        float vf[8];
        unpack8(*reinterpret_cast<const uint4*>(base + j * ld + 2 * Da), vf);
        const float p = sc[j];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, vf[i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = across_groups_sum<LPR>(acc[i]);
    if (lane < LPR) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[warp * HS + sub * 8 + i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < HS) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < kClsWarps; ++w) o += red[w * HS + threadIdx.x];
        out[static_cast<size_t>(n) * T * Da + static_cast<size_t>(h) * HS + threadIdx.x] = __float2bfloat16(o / sum);
    }
    if (threadIdx.x == 0) lse[(static_cast<size_t>(n) * heads + h) * T] = mx + logf(sum);
}

// Backward of the above for head h of cutout n.  With p_j = exp(q.k_j - lse), dp_j = dO.v_j, delta = dO.o:
//   dV_j = p_j dO,  dK_j = ds_j q,  dQ_0 = sum_j ds_j k_j,  ds_j = p_j (dp_j - delta);  dQ rows j > 0 are zero.
// Every (row, q|k|v, head) slice of d_qkv is written (the following dgrad GEMM reads all of it).
template <int HS>
__global__ void __launch_bounds__(kClsThreads)
attn_cls_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ d_out,
                    const float* __restrict__ lse, bf16* __restrict__ d_qkv, int T, int heads) {
    grid_dep_launch();
    constexpr int LPR = HS / 8, KPW = 32 / LPR, KSTEP = KPW * kClsWarps;
    __shared__ float red[kClsWarps][HS];
    const int h = blockIdx.x, n = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const size_t Da = static_cast<size_t>(heads) * HS, ld = 3 * Da;
    const size_t row0 = static_cast<size_t>(n) * T;
    const size_t col = static_cast<size_t>(h) * HS + sub * 8;
    const bf16* base = qkv + row0 * ld + col;
    bf16* dbase = d_qkv + row0 * ld + col;
    float qf[8], of[8], dof[8];
    unpack8(*reinterpret_cast<const uint4*>(base), qf);
    unpack8(*reinterpret_cast<const uint4*>(out + row0 * Da + col), of);
    unpack8(*reinterpret_cast<const uint4*>(d_out + row0 * Da + col), dof);
    const float delta = group_sum<LPR>(dot8(dof, of));
    const float L = lse[(static_cast<size_t>(n) * heads + h) * T];
    float dq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = warp * KPW; j0 < T; j0 += KSTEP) {  // warp-uniform trip count: the group sums shuffle across the warp
        const int j = j0 + lane / LPR;
        const bool ok = j < T;
        float kf[8], vf[8];
        unpack8(ok ? *reinterpret_cast<const uint4*>(base + j * ld + Da) : make_uint4(0u, 0u, 0u, 0u), kf);
        unpack8(ok ? *reinterpret_cast<const uint4*>(base + j * ld + 2 * Da) : make_uint4(0u, 0u, 0u, 0u), vf);
        const float s = group_sum<LPR>(dot8(qf, kf));
        const float dp = group_sum<LPR>(dot8(dof, vf));
        if (!ok) continue;
        const float p = __expf(s - L);
        const float ds = p * (dp - delta);
        *reinterpret_cast<uint4*>(dbase + j * ld + 2 * Da) =
            make_uint4(pack_bf16(p * dof[0], p * dof[1]), pack_bf16(p * dof[2], p * dof[3]),
                       pack_bf16(p * dof[4], p * dof[5]), pack_bf16(p * dof[6], p * dof[7]));
        *reinterpret_cast<uint4*>(dbase + j * ld + Da) =
            make_uint4(pack_bf16(ds * qf[0], ds * qf[1]), pack_bf16(ds * qf[2], ds * qf[3]),
                       pack_bf16(ds * qf[4], ds * qf[5]), pack_bf16(ds * qf[6], ds * qf[7]));
        if (j != 0) *reinterpret_cast<uint4*>(dbase + j * ld) = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < 8; ++i) dq[i] = fmaf(ds, kf[i], dq[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i] = across_groups_sum<LPR>(dq[i]);
    if (lane < LPR) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[warp][sub * 8 + i] = dq[i];
    }
    __syncthreads();
    if (threadIdx.x < HS) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kClsWarps; ++w) v += red[w][threadIdx.x];
        d_qkv[row0 * ld + static_cast<size_t>(h) * HS + threadIdx.x] = __float2bfloat16(v);
    }
}

int check_cls(const char* who, int n, int T, int heads, int head_stride) {
    PCG_CHECK_ARG(n > 0 && T > 0 && heads > 0, "%s: bad sizes n=%d T=%d heads=%d", who, n, T, heads);
    PCG_CHECK_ARG(head_stride == 64 || head_stride == 128, "%s: head stride %d is not 64 or 128", who, head_stride);
    PCG_CHECK_ARG(T <= 8192, "%s: T=%d exceeds 8192", who, T);
    PCG_CHECK_ARG(n <= 65535 && heads <= 65535, "%s: n and heads must be <= 65535 (grid = heads x n)", who);
    return 0;
}

}  // namespace
}  // namespace pcg

using namespace pcg;

extern "C" int pcg_attn_cls_fwd(const void* qkv, void* out, float* lse, int n, int T, int heads, int head_stride,
                                void* stream) {
    if (int rc = check_cls("pcg_attn_cls_fwd", n, T, heads, head_stride)) return rc;
    PCG_CHECK_ARG(qkv && out && lse, "pcg_attn_cls_fwd: null buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_FWD, 4.0 * n * heads * T * head_stride, s);
    const size_t smem = (((T + 3) & ~3) + kClsWarps * head_stride + 2 * kClsWarps) * sizeof(float);
    const dim3 grid(heads, n);
    if (head_stride == 64)
        attn_cls_fwd_kernel<64><<<grid, kClsThreads, smem, s>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, T, heads);
    else
        attn_cls_fwd_kernel<128><<<grid, kClsThreads, smem, s>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, T, heads);
    PCG_LAUNCH_CHECK("attn_cls_fwd_kernel");
    return 0;
}

extern "C" int pcg_attn_cls_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* d_qkv,
                                int n, int T, int heads, int head_stride, void* stream) {
    if (int rc = check_cls("pcg_attn_cls_bwd", n, T, heads, head_stride)) return rc;
    PCG_CHECK_ARG(qkv && out && d_out && lse && d_qkv, "pcg_attn_cls_bwd: null buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_BWD, 8.0 * n * heads * T * head_stride, s);
    const dim3 grid(heads, n);
    if (head_stride == 64)
        attn_cls_bwd_kernel<64><<<grid, kClsThreads, 0, s>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(out),
                                                             static_cast<const bf16*>(d_out), lse, static_cast<bf16*>(d_qkv), T, heads);
    else
        attn_cls_bwd_kernel<128><<<grid, kClsThreads, 0, s>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(out),
                                                              static_cast<const bf16*>(d_out), lse, static_cast<bf16*>(d_qkv), T, heads);
    PCG_LAUNCH_CHECK("attn_cls_bwd_kernel");
    return 0;
}
