// S0+S1+S2: random-cutout gather + separable antialiased resize (ResizeRight semantics) + CLIP normalisation,
// written straight into the patch-major bf16 operand of the patch-embedding GEMM; and its backward, which
// scatter-adds the gradient of every cutout into the shared fp32 image gradient.
//
// The crop boxes (S0) and the per-size tap tables (first input index, weights; S1) are produced on the host with
// the reference's own fp32 arithmetic, so indices are bit-exact by construction; this kernel only applies them.
// One CTA = (cutout, channel, RB output rows): pass 1 resamples vertically (reads coalesced along x, the source
// image is L2 resident), the [RB, in_w] intermediate lives in shared memory, pass 2 resamples horizontally,
// normalises and stores.  HBM-bound: algorithmic bytes = 3*R*R*2 B written per cutout (+ crop footprint read).
//
// Replaces resize() + Normalize in perceptor/models/open_clip.py:109-118
// (perceptor/transforms/resize/resize_right.py:34-189, apply_weights :288-318, zero padding :44).
#include "pcg_common.cuh"
#include "pcg_ptx.cuh"

namespace pcg {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSmemBudget = 64 * 1024;   // per CTA target, so that 3 CTAs share an SM
constexpr int kSmemMax = 200 * 1024;

struct SamplerParams {
    const float* images;
    int B, H, W;
    const int32_t* cuts;
    const int32_t* desc;
    const int32_t* left;
    const float* weight;
    const int32_t* inv;
    int R, patch, grid, kpad, RB;
    float mean[3], inv_std[3];
};

struct Cut {
    int b, y0, x0, sh, sw;
    int taps_v, taps_h;
    const int32_t *left_v, *left_h, *inv_h;
    const float *w_v, *w_h;
};

__device__ __forceinline__ Cut load_cut(const SamplerParams& p, int n) {
    const int32_t* c = p.cuts + n * PCG_CUT_STRIDE;
    Cut k;
    k.b = c[0]; k.y0 = c[1]; k.x0 = c[2]; k.sh = c[3]; k.sw = c[4];
    const int32_t* dv = p.desc + c[5] * 8;
    const int32_t* dh = p.desc + c[6] * 8;
    k.taps_v = dv[0]; k.left_v = p.left + dv[1]; k.w_v = p.weight + dv[2];
    k.taps_h = dh[0]; k.left_h = p.left + dh[1]; k.w_h = p.weight + dh[2]; k.inv_h = p.inv + dh[3];
    return k;
}

__global__ void __launch_bounds__(kThreads) sampler_fwd_kernel(const SamplerParams p, bf16* __restrict__ patches,
                                                               float* __restrict__ out_f32) {
    extern __shared__ float tmp[];  // [RB][sw]
    const int n = blockIdx.z, ch = blockIdx.y, r0 = blockIdx.x * p.RB;
    const int nr = min(p.RB, p.R - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Cut k = load_cut(p, n);
    const float* src = p.images + ((static_cast<size_t>(k.b) * 3 + ch) * p.H + k.y0) * p.W + k.x0;

    // pass 1: vertical.  tmp[r][x] = sum_t w_v[o][t] * crop[left_v[o] + t][x]   (rows outside the crop are zero)
    for (int r = warp; r < nr; r += kWarps) {
        const int o = r0 + r;
        const int left = __ldg(k.left_v + o);
        const float* wrow = k.w_v + static_cast<size_t>(o) * k.taps_v;
        const int t_lo = max(0, -left), t_hi = min(k.taps_v, k.sh - left);
        for (int x = lane; x < k.sw; x += 32) {
            float acc = 0.f;
            for (int t = t_lo; t < t_hi; ++t) acc = fmaf(__ldg(wrow + t), __ldg(src + static_cast<size_t>(left + t) * p.W + x), acc);
            tmp[r * k.sw + x] = acc;
        }
    }
    __syncthreads();
    // pass 2: horizontal + normalise + store
    const float mean = p.mean[ch], inv_std = p.inv_std[ch];
    const int pp = p.patch * p.patch;
    for (int r = warp; r < nr; r += kWarps) {
        const int o = r0 + r;
        const int gy = o / p.patch, py = o - gy * p.patch;
        const float* trow = tmp + r * k.sw;
        for (int c = lane; c < p.R; c += 32) {
            const int left = __ldg(k.left_h + c);
            const float* wrow = k.w_h + static_cast<size_t>(c) * k.taps_h;
            const int t_lo = max(0, -left), t_hi = min(k.taps_h, k.sw - left);
            float acc = 0.f;
            for (int t = t_lo; t < t_hi; ++t) acc = fmaf(__ldg(wrow + t), trow[left + t], acc);
            const float v = (acc - mean) * inv_std;
            if (out_f32 != nullptr) out_f32[((static_cast<size_t>(n) * 3 + ch) * p.R + o) * p.R + c] = v;
            if (patches != nullptr) {
                const int gx = c / p.patch, px = c - gx * p.patch;
                const size_t row = (static_cast<size_t>(n) * p.grid + gy) * p.grid + gx;
                patches[row * p.kpad + ch * pp + py * p.patch + px] = __float2bfloat16(v);
            }
        }
    }
}

// backward: same tiling.  d_out tile -> (horizontal^T) -> d_tmp[RB][sw] in smem -> (vertical^T) -> atomicAdd.
__global__ void __launch_bounds__(kThreads) sampler_bwd_kernel(const SamplerParams p, const bf16* __restrict__ d_patches,
                                                               const float* __restrict__ d_out_f32,
                                                               float* __restrict__ d_images) {
    extern __shared__ float smf[];
    const int n = blockIdx.z, ch = blockIdx.y, r0 = blockIdx.x * p.RB;
    const int nr = min(p.RB, p.R - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Cut k = load_cut(p, n);
    float* d_tmp = smf;                     // [RB][sw]
    float* dout = smf + p.RB * k.sw;        // [RB][R]
    const float inv_std = p.inv_std[ch];
    const int pp = p.patch * p.patch;

    for (int r = warp; r < nr; r += kWarps) {
        const int o = r0 + r;
        const int gy = o / p.patch, py = o - gy * p.patch;
        for (int c = lane; c < p.R; c += 32) {
            float g;
            if (d_patches != nullptr) {
                const int gx = c / p.patch, px = c - gx * p.patch;
                const size_t row = (static_cast<size_t>(n) * p.grid + gy) * p.grid + gx;
                g = __bfloat162float(d_patches[row * p.kpad + ch * pp + py * p.patch + px]);
            } else {
                g = d_out_f32[((static_cast<size_t>(n) * 3 + ch) * p.R + o) * p.R + c];
            }
            dout[r * p.R + c] = g * inv_std;
        }
    }
    __syncthreads();
    // horizontal transpose: d_tmp[r][x] = sum_{c in inv_h[x]} w_h[c][x - left_h[c]] * dout[r][c]
    for (int r = warp; r < nr; r += kWarps) {
        const float* drow = dout + r * p.R;
        for (int x = lane; x < k.sw; x += 32) {
            const int c_lo = __ldg(k.inv_h + 2 * x), c_hi = __ldg(k.inv_h + 2 * x + 1);
            float acc = 0.f;
            for (int c = c_lo; c < c_hi; ++c) {
                const int t = x - __ldg(k.left_h + c);
                if (t >= 0 && t < k.taps_h) acc = fmaf(__ldg(k.w_h + static_cast<size_t>(c) * k.taps_h + t), drow[c], acc);
            }
            d_tmp[r * k.sw + x] = acc;
        }
    }
    __syncthreads();
    // vertical transpose, reduced over the tile's rows before touching global memory
    const int y_lo = max(0, __ldg(k.left_v + r0));
    const int y_hi = min(k.sh, __ldg(k.left_v + r0 + nr - 1) + k.taps_v);
    float* dst = d_images + ((static_cast<size_t>(k.b) * 3 + ch) * p.H + k.y0) * p.W + k.x0;
    for (int y = y_lo + warp; y < y_hi; y += kWarps) {
        for (int x = lane; x < k.sw; x += 32) {
            float acc = 0.f;
            for (int r = 0; r < nr; ++r) {
                const int o = r0 + r;
                const int t = y - __ldg(k.left_v + o);
                if (t >= 0 && t < k.taps_v) acc = fmaf(__ldg(k.w_v + static_cast<size_t>(o) * k.taps_v + t), d_tmp[r * k.sw + x], acc);
            }
            atomicAdd(dst + static_cast<size_t>(y) * p.W + x, acc);
        }
    }
}

int pick_rb(int row_floats, int R) {
    int rb = 32;
    while (rb > 1 && static_cast<size_t>(rb) * row_floats * sizeof(float) > kSmemBudget) rb >>= 1;
    if (rb > R) rb = R;
    return rb;
}

int fill_params(SamplerParams& p, const float* images, int B, int H, int W, const int32_t* cuts,
                const pcg_resize_tables* tabs, int R, int patch, int kpad, const float* mean_host,
                const float* std_host) {
    p.images = images; p.B = B; p.H = H; p.W = W; p.cuts = cuts;
    p.desc = tabs->desc; p.left = tabs->left; p.weight = tabs->weight; p.inv = tabs->inv;
    p.R = R; p.patch = patch; p.grid = R / patch; p.kpad = kpad;
    for (int i = 0; i < 3; ++i) {
        p.mean[i] = mean_host ? mean_host[i] : 0.f;
        p.inv_std[i] = std_host ? 1.0f / std_host[i] : 1.f;
    }
    return 0;
}

}  // namespace
}  // namespace pcg

using namespace pcg;

extern "C" int pcg_sampler_fwd(const float* images, int B, int H, int W, const int32_t* cuts, int n_cut,
                               const pcg_resize_tables* tabs, int R, int patch, int kpad, const float* mean_host,
                               const float* std_host, void* patches_bf16, float* out_f32, int max_in_w, void* stream) {
    PCG_CHECK_ARG(images && cuts && tabs && tabs->desc && tabs->left && tabs->weight, "pcg_sampler_fwd: null pointer");
    PCG_CHECK_ARG(patches_bf16 || out_f32, "pcg_sampler_fwd: no output requested");
    PCG_CHECK_ARG(n_cut > 0 && n_cut <= 65535 && R > 0 && patch > 0 && R % patch == 0 && max_in_w > 0,
                  "pcg_sampler_fwd: bad shape n_cut=%d R=%d patch=%d max_in_w=%d", n_cut, R, patch, max_in_w);
    PCG_CHECK_ARG(patches_bf16 == nullptr || kpad >= 3 * patch * patch, "pcg_sampler_fwd: kpad %d < 3*p*p", kpad);
    SamplerParams p;
    fill_params(p, images, B, H, W, cuts, tabs, R, patch, kpad, mean_host, std_host);
    p.RB = pick_rb(max_in_w, R);
    const size_t smem = static_cast<size_t>(p.RB) * max_in_w * sizeof(float);
    PCG_CHECK_ARG(smem <= kSmemMax, "pcg_sampler_fwd: crop width %d too large", max_in_w);
    static bool configured = false;
    if (!configured) {
        PCG_CUDA(cudaFuncSetAttribute(sampler_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        configured = true;
    }
    const dim3 grid(ceil_div(R, p.RB), 3, n_cut);
    ProfileScope prof(PCG_PROF_SAMPLER_FWD, 6.0 * n_cut * R * R, static_cast<cudaStream_t>(stream));
    sampler_fwd_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p, static_cast<bf16*>(patches_bf16),
                                                                                    out_f32);
    PCG_LAUNCH_CHECK("sampler_fwd_kernel");
    return 0;
}

extern "C" int pcg_sampler_bwd(const void* d_patches_bf16, const float* d_out_f32, int B, int H, int W,
                               const int32_t* cuts, int n_cut, const pcg_resize_tables* tabs, int R, int patch, int kpad,
                               const float* std_host, float* d_images, int max_in_w, void* stream) {
    PCG_CHECK_ARG(cuts && tabs && tabs->desc && tabs->left && tabs->weight && tabs->inv && d_images,
                  "pcg_sampler_bwd: null pointer");
    PCG_CHECK_ARG(d_patches_bf16 || d_out_f32, "pcg_sampler_bwd: no input gradient");
    PCG_CHECK_ARG(n_cut > 0 && n_cut <= 65535 && R > 0 && patch > 0 && R % patch == 0 && max_in_w > 0,
                  "pcg_sampler_bwd: bad shape n_cut=%d R=%d patch=%d max_in_w=%d", n_cut, R, patch, max_in_w);
    SamplerParams p;
    fill_params(p, nullptr, B, H, W, cuts, tabs, R, patch, kpad, nullptr, std_host);
    p.RB = pick_rb(max_in_w + R, R);
    const size_t smem = static_cast<size_t>(p.RB) * (max_in_w + R) * sizeof(float);
    PCG_CHECK_ARG(smem <= kSmemMax, "pcg_sampler_bwd: crop width %d too large", max_in_w);
    static bool configured = false;
    if (!configured) {
        PCG_CUDA(cudaFuncSetAttribute(sampler_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        configured = true;
    }
    const dim3 grid(ceil_div(R, p.RB), 3, n_cut);
    ProfileScope prof(PCG_PROF_SAMPLER_BWD, 6.0 * n_cut * R * R, static_cast<cudaStream_t>(stream));
    sampler_bwd_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        p, static_cast<const bf16*>(d_patches_bf16), d_out_f32, d_images);
    PCG_LAUNCH_CHECK("sampler_bwd_kernel");
    return 0;
}
