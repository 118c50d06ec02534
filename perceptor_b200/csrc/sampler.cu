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
// Two kernel generations share that tiling.  The `_vec` kernels are the production path (image rows 16-byte
// aligned, R and patch even): every inner loop is vectorised along the dimension that is NOT being resampled, so
// one tap weight serves four multiply-adds -- pass 1 reads the image as float4 from a crop origin rounded down
// to a multiple of four columns, pass 2 keeps four rows per thread and reads its weights from a transposed copy
// in shared memory; the backward mirrors this and ends in 16-byte vector reductions (red.global.add.v4.f32) into
// the image gradient.  The scalar kernels remain for shapes the vector path cannot take (W % 4 != 0, odd R).
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
    int taps_cap;  // vector kernels: rows of the transposed weight copy the launch sized shared memory for
    int RS;  // vector kernels: row stride of the shared-memory intermediate, floats (multiple of 4)
    float mean[3], inv_std[3];
};

// *addr[0..3] += v as one 16-byte reduction (sm_90+)
__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

struct Cut {
    int b, y0, x0, sh, sw;
    int taps_v, taps_h;
    const int32_t *left_v, *left_h, *inv_v, *inv_h;
    const float *w_v, *w_h;
};

__device__ __forceinline__ Cut load_cut(const SamplerParams& p, int n) {
    const int32_t* c = p.cuts + n * PCG_CUT_STRIDE;
    Cut k;
    k.b = c[0]; k.y0 = c[1]; k.x0 = c[2]; k.sh = c[3]; k.sw = c[4];
    const int32_t* dv = p.desc + c[5] * 8;
    const int32_t* dh = p.desc + c[6] * 8;
    k.taps_v = dv[0]; k.left_v = p.left + dv[1]; k.w_v = p.weight + dv[2]; k.inv_v = p.inv + dv[3];
    k.taps_h = dh[0]; k.left_h = p.left + dh[1]; k.w_h = p.weight + dh[2]; k.inv_h = p.inv + dh[3];
    return k;
}

__global__ void __launch_bounds__(kThreads) sampler_fwd_kernel(const SamplerParams p, bf16* __restrict__ patches,
                                                               float* __restrict__ out_f32) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
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


// ---------------------------------------------------------------------------------------------------------
// vectorised kernels
// ---------------------------------------------------------------------------------------------------------
// shared memory: tmp [RB][RS] floats (RS % 4 == 0; column index = crop x + (x0 & 3)), then the horizontal tap
// weights transposed, wT[t * (R + 1) + c] (the +1 keeps (t + 1, c) and (t, c + 1) in different banks)
__global__ void __launch_bounds__(kThreads) sampler_fwd_vec_kernel(const SamplerParams p, bf16* __restrict__ patches,
                                                                   float* __restrict__ out_f32) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    extern __shared__ float4 smem4[];
    float* tmp = reinterpret_cast<float*>(smem4);
    const int RS = p.RS;
    float* wT = tmp + p.RB * RS;
    const int n = blockIdx.z, ch = blockIdx.y, r0 = blockIdx.x * p.RB;
    const int nr = min(p.RB, p.R - r0);
    const Cut k = load_cut(p, n);
    const int xa = k.x0 & 3;
    const int nxg = (k.sw + xa + 3) >> 2;
    const int w4 = p.W >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(
        p.images + ((static_cast<size_t>(k.b) * 3 + ch) * p.H + k.y0) * p.W + (k.x0 - xa));

    if (k.taps_h > p.taps_cap) __trap();  // the host bound on the tap count is wrong: never continue silently
    for (int i = threadIdx.x; i < p.R * k.taps_h; i += kThreads) {
        const int c = i / k.taps_h, t = i - c * k.taps_h;
        wT[t * (p.R + 1) + c] = __ldg(k.w_h + i);
    }
    // pass 1: vertical, four columns per thread.  Columns left of the crop (the xa alignment columns) and right
    // of it hold neighbouring pixels; pass 2 never reads them.
    for (int item = threadIdx.x; item < nr * nxg; item += kThreads) {
        const int r = item / nxg, xg = item - r * nxg;
        const int o = r0 + r;
        const int left = __ldg(k.left_v + o);
        const float* wrow = k.w_v + static_cast<size_t>(o) * k.taps_v;
        const int t_lo = max(0, -left), t_hi = min(k.taps_v, k.sh - left);
        const float4* s = src4 + static_cast<size_t>(left + t_lo) * w4 + xg;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int t = t_lo; t < t_hi; ++t, s += w4) {
            const float w = __ldg(wrow + t);
            const float4 v = __ldg(s);
            acc.x = fmaf(w, v.x, acc.x);
            acc.y = fmaf(w, v.y, acc.y);
            acc.z = fmaf(w, v.z, acc.z);
            acc.w = fmaf(w, v.w, acc.w);
        }
        smem4[(r * RS >> 2) + xg] = acc;
    }
    __syncthreads();
    // pass 2: horizontal, four rows per thread, + normalise + store
    const float mean = p.mean[ch], inv_std = p.inv_std[ch];
    const int pp = p.patch * p.patch;
    const int nrg = (nr + 3) >> 2;
    for (int item = threadIdx.x; item < nrg * p.R; item += kThreads) {
        const int rg = item / p.R, c = item - rg * p.R;
        const int left = __ldg(k.left_h + c);
        const int t_lo = max(0, -left), t_hi = min(k.taps_h, k.sw - left);
        const float* t0 = tmp + (4 * rg) * RS + xa + left;
        const float* wc = wT + c;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 2
        for (int t = t_lo; t < t_hi; ++t) {
            const float w = wc[t * (p.R + 1)];
            a0 = fmaf(w, t0[t], a0);
            a1 = fmaf(w, t0[RS + t], a1);
            a2 = fmaf(w, t0[2 * RS + t], a2);
            a3 = fmaf(w, t0[3 * RS + t], a3);
        }
        const float acc[4] = {a0, a1, a2, a3};
        const int gx = c / p.patch, px = c - gx * p.patch;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = r0 + 4 * rg + j;
            if (4 * rg + j >= nr) break;
            const float v = (acc[j] - mean) * inv_std;
            if (out_f32 != nullptr) out_f32[((static_cast<size_t>(n) * 3 + ch) * p.R + o) * p.R + c] = v;
            if (patches != nullptr) {
                const int gy = o / p.patch, py = o - gy * p.patch;
                const size_t row = (static_cast<size_t>(n) * p.grid + gy) * p.grid + gx;
                patches[row * p.kpad + ch * pp + py * p.patch + px] = __float2bfloat16(v);
            }
        }
    }
}

// backward.  shared memory: dout [RB][R], d_tmp [RB][RS], wT as above.
__global__ void __launch_bounds__(kThreads) sampler_bwd_vec_kernel(const SamplerParams p, const bf16* __restrict__ d_patches,
                                                                   const float* __restrict__ d_out_f32,
                                                                   float* __restrict__ d_images) {
    extern __shared__ float4 smem4[];
    const int RS = p.RS;
    float* d_tmp = reinterpret_cast<float*>(smem4);  // [RB][RS]
    float* dout = d_tmp + p.RB * RS;                 // [RB][R]
    float* wT = dout + p.RB * p.R;
    const int n = blockIdx.z, ch = blockIdx.y, r0 = blockIdx.x * p.RB;
    const int nr = min(p.RB, p.R - r0);
    const Cut k = load_cut(p, n);
    const int xa = k.x0 & 3;
    const int nxg = (k.sw + xa + 3) >> 2;
    const float inv_std = p.inv_std[ch];
    const int pp = p.patch * p.patch;
    const int half_r = p.R >> 1;

    if (k.taps_h > p.taps_cap) __trap();  // the host bound on the tap count is wrong: never continue silently
    for (int i = threadIdx.x; i < p.R * k.taps_h; i += kThreads) {
        const int c = i / k.taps_h, t = i - c * k.taps_h;
        wT[t * (p.R + 1) + c] = __ldg(k.w_h + i);
    }
    // incoming gradient tile, two columns per thread (R and patch are even: a pair never straddles a patch);
    // rows past the end of the image are zero so that the four-row groups below need no row guard
    for (int item = threadIdx.x; item < p.RB * half_r; item += kThreads) {
        const int r = item / half_r, c = (item - r * half_r) * 2;
        float g0 = 0.f, g1 = 0.f;
        if (r < nr) {
            const int o = r0 + r;
            if (d_patches != nullptr) {
                const int gy = o / p.patch, py = o - gy * p.patch;
                const int gx = c / p.patch, px = c - gx * p.patch;
                const size_t row = (static_cast<size_t>(n) * p.grid + gy) * p.grid + gx;
                const __nv_bfloat162 g2 =
                    *reinterpret_cast<const __nv_bfloat162*>(d_patches + row * p.kpad + ch * pp + py * p.patch + px);
                g0 = __low2float(g2), g1 = __high2float(g2);
            } else {
                const float2 g2 = *reinterpret_cast<const float2*>(
                    d_out_f32 + ((static_cast<size_t>(n) * 3 + ch) * p.R + o) * p.R + c);
                g0 = g2.x, g1 = g2.y;
            }
        }
        *reinterpret_cast<float2*>(dout + r * p.R + c) = make_float2(g0 * inv_std, g1 * inv_std);
    }
    __syncthreads();
    // horizontal transpose, four rows per thread: d_tmp[r][xa + x] = sum_c w_h[c][x - left_h[c]] * dout[r][c]
    const int nrg = p.RB >> 2;
    for (int item = threadIdx.x; item < nrg * nxg * 4; item += kThreads) {
        const int rg = item / (nxg * 4), xs = item - rg * (nxg * 4);
        const int x = xs - xa;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (x >= 0 && x < k.sw) {
            const int c_lo = __ldg(k.inv_h + 2 * x), c_hi = __ldg(k.inv_h + 2 * x + 1);
            const float* d0 = dout + (4 * rg) * p.R;
            for (int c = c_lo; c < c_hi; ++c) {
                const int t = x - __ldg(k.left_h + c);
                if (t >= 0 && t < k.taps_h) {
                    const float w = wT[t * (p.R + 1) + c];
                    a0 = fmaf(w, d0[c], a0);
                    a1 = fmaf(w, d0[p.R + c], a1);
                    a2 = fmaf(w, d0[2 * p.R + c], a2);
                    a3 = fmaf(w, d0[3 * p.R + c], a3);
                }
            }
        }
        float* dt = d_tmp + (4 * rg) * RS + xs;
        dt[0] = a0, dt[RS] = a1, dt[2 * RS] = a2, dt[3 * RS] = a3;
    }
    __syncthreads();
    // vertical transpose over the tile's rows, four columns per thread, then one vector reduction per 16 bytes
    const int y_lo = max(0, __ldg(k.left_v + r0));
    const int y_hi = min(k.sh, __ldg(k.left_v + r0 + nr - 1) + k.taps_v);
    float* dst = d_images + ((static_cast<size_t>(k.b) * 3 + ch) * p.H + k.y0) * p.W + (k.x0 - xa);
    for (int item = threadIdx.x; item < (y_hi - y_lo) * nxg; item += kThreads) {
        const int yy = item / nxg, xg = item - yy * nxg;
        const int y = y_lo + yy;
        const int r_lo = max(__ldg(k.inv_v + 2 * y), r0), r_hi = min(__ldg(k.inv_v + 2 * y + 1), r0 + nr);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int o = r_lo; o < r_hi; ++o) {
            const int t = y - __ldg(k.left_v + o);
            if (t >= 0 && t < k.taps_v) {
                const float w = __ldg(k.w_v + static_cast<size_t>(o) * k.taps_v + t);
                const float4 v = smem4[((o - r0) * RS >> 2) + xg];
                acc.x = fmaf(w, v.x, acc.x);
                acc.y = fmaf(w, v.y, acc.y);
                acc.z = fmaf(w, v.z, acc.z);
                acc.w = fmaf(w, v.w, acc.w);
            }
        }
        float* d = dst + static_cast<size_t>(y) * p.W + 4 * xg;
        const int x = 4 * xg - xa;
        if (x >= 0 && x + 3 < k.sw) {
            red_add_v4(d, acc);
        } else {
            if (x >= 0 && x < k.sw) atomicAdd(d, acc.x);
            if (x + 1 >= 0 && x + 1 < k.sw) atomicAdd(d + 1, acc.y);
            if (x + 2 >= 0 && x + 2 < k.sw) atomicAdd(d + 2, acc.z);
            if (x + 3 >= 0 && x + 3 < k.sw) atomicAdd(d + 3, acc.w);
        }
    }
}

// the vector path needs 16-byte aligned image rows and even R / patch
bool vec_ok(const void* images, int W, int R, int patch) {
    return (reinterpret_cast<uintptr_t>(images) & 15u) == 0 && (W & 3) == 0 && (R & 1) == 0 && (patch & 1) == 0;
}
int vec_row_stride(int max_in_w) { return (max_in_w + 3 + 3) & ~3; }
// upper bound of the horizontal tap count of any table with in_size <= max_in_w: lanczos3 stretched by in/out
// has ceil(6 in / out) taps (resize_right.py:426-436), bicubic 4, the identity 1
int vec_taps_cap(int max_in_w, int R) { return max(4, (6 * max_in_w + R - 1) / R + 1); }
bool g_force_scalar = false;
// rows per tile: as many as fit the per-CTA budget (several CTAs per SM), a multiple of four
int pick_rb_vec(size_t bytes_per_row, size_t fixed_bytes, int R) {
    int rb = 16;
    while (rb > 4 && rb * bytes_per_row + fixed_bytes > static_cast<size_t>(kSmemBudget)) rb >>= 1;
    (void)R;
    return rb;
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

extern "C" int pcg_sampler_set_scalar(int on) {  // test hook: force the scalar kernels
    g_force_scalar = on != 0;
    return 0;
}

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
    static PerDeviceOnce configured;
    PCG_ONCE_PER_DEVICE(configured, PCG_CUDA(cudaFuncSetAttribute(sampler_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)); PCG_CUDA(cudaFuncSetAttribute(sampler_fwd_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)));
    ProfileScope prof(PCG_PROF_SAMPLER_FWD, 6.0 * n_cut * R * R, static_cast<cudaStream_t>(stream));
    if (!g_force_scalar && vec_ok(images, W, R, patch)) {
        p.RS = vec_row_stride(max_in_w);
        p.taps_cap = vec_taps_cap(max_in_w, R);
        const size_t fixed = static_cast<size_t>(p.taps_cap) * (R + 1) * sizeof(float);
        p.RB = pick_rb_vec(p.RS * sizeof(float), fixed, R);
        const size_t smem_vec = static_cast<size_t>(p.RB) * p.RS * sizeof(float) + fixed;
        if (smem_vec <= kSmemMax) {
            const dim3 grid_vec(ceil_div(R, p.RB), 3, n_cut);
            sampler_fwd_vec_kernel<<<grid_vec, kThreads, smem_vec, static_cast<cudaStream_t>(stream)>>>(
                p, static_cast<bf16*>(patches_bf16), out_f32);
            PCG_LAUNCH_CHECK("sampler_fwd_vec_kernel");
            return 0;
        }
    }
    p.RB = pick_rb(max_in_w, R);
    const size_t smem = static_cast<size_t>(p.RB) * max_in_w * sizeof(float);
    PCG_CHECK_ARG(smem <= kSmemMax, "pcg_sampler_fwd: crop width %d too large", max_in_w);
    const dim3 grid(ceil_div(R, p.RB), 3, n_cut);
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
    static PerDeviceOnce configured;
    PCG_ONCE_PER_DEVICE(configured, PCG_CUDA(cudaFuncSetAttribute(sampler_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)); PCG_CUDA(cudaFuncSetAttribute(sampler_bwd_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)));
    ProfileScope prof(PCG_PROF_SAMPLER_BWD, 6.0 * n_cut * R * R, static_cast<cudaStream_t>(stream));
    if (!g_force_scalar && vec_ok(d_images, W, R, patch) &&
        (d_out_f32 == nullptr || (reinterpret_cast<uintptr_t>(d_out_f32) & 7u) == 0)) {
        p.RS = vec_row_stride(max_in_w);
        p.taps_cap = vec_taps_cap(max_in_w, R);
        const size_t fixed = static_cast<size_t>(p.taps_cap) * (R + 1) * sizeof(float);
        p.RB = pick_rb_vec((p.RS + R) * sizeof(float), fixed, R);
        const size_t smem_vec = static_cast<size_t>(p.RB) * (p.RS + R) * sizeof(float) + fixed;
        if (smem_vec <= kSmemMax) {
            const dim3 grid_vec(ceil_div(R, p.RB), 3, n_cut);
            sampler_bwd_vec_kernel<<<grid_vec, kThreads, smem_vec, static_cast<cudaStream_t>(stream)>>>(
                p, static_cast<const bf16*>(d_patches_bf16), d_out_f32, d_images);
            PCG_LAUNCH_CHECK("sampler_bwd_vec_kernel");
            return 0;
        }
    }
    p.RB = pick_rb(max_in_w + R, R);
    const size_t smem = static_cast<size_t>(p.RB) * (max_in_w + R) * sizeof(float);
    PCG_CHECK_ARG(smem <= kSmemMax, "pcg_sampler_bwd: crop width %d too large", max_in_w);
    const dim3 grid(ceil_div(R, p.RB), 3, n_cut);
    sampler_bwd_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        p, static_cast<const bf16*>(d_patches_bf16), d_out_f32, d_images);
    PCG_LAUNCH_CHECK("sampler_bwd_kernel");
    return 0;
}
