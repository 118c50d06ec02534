// Guided-sampling glue either side of the guidance loss: the per-sample affine maps of a
// velocity-diffusion step and the straight-through clamp, each one fused, vectorised HBM-bound kernel instead of
// the reference's chains of eager elementwise ops.
//
// Everything a v-objective step does to an image batch is affine per sample n with coefficients that depend only
// on the timesteps (alpha = cos(t pi/2), sigma = sin(t pi/2), perceptor/models/velocity_diffusion/utils.py:47-50):
//   denoised_images = decode(encode(x) alpha - v sigma)            = alpha x - (sigma/2) v + (1 - alpha)/2
//   step(to_ts, eta=0) = decode(den_xs alpha' + eps sigma')        = A x + (alpha sigma' - sigma alpha')/2 v + (1 - A)/2,
//                                                                    A = alpha alpha' + sigma sigma'
//   guided(g)       = v + scale sigma clamp(g, -c, c) / c
// (perceptor/models/velocity_diffusion/predictions.py:50-66, 68-105, 148-155; encode/decode diffusion_space.py:1-6)
// so one kernel  out[n, i] = a[n] x[n, i] + b[n] clamp(y[n, i], -lim, lim) + c[n]  covers the forward maps and, with
// y = null, their backward.  clamp_with_grad's backward (perceptor/transforms/clamp_with_grad.py:17-23) is the
// second kernel.  Algorithmic bytes: 4 B per element per tensor touched.
#include <algorithm>

#include "pcg_common.cuh"

namespace pcg {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float clampf(float v, float lim) { return fminf(fmaxf(v, -lim), lim); }

template <bool kHasY, bool kVec>
__global__ void __launch_bounds__(kThreads) affine2_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           const float* __restrict__ coef, float* __restrict__ out,
                                                           int n_samples, size_t per, float lim) {
    // coef = [3][n_samples]: a, b, c.  grid.y = sample, grid.x strides over the sample's elements
    const int n = blockIdx.y;
    const float a = __ldg(coef + n), b = __ldg(coef + n_samples + n), c = __ldg(coef + 2 * n_samples + n);
    const size_t base = static_cast<size_t>(n) * per;
    if (kVec) {
        const size_t per4 = per >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(x + base);
        const float4* y4 = reinterpret_cast<const float4*>(y + base);
        float4* o4 = reinterpret_cast<float4*>(out + base);
        for (size_t i = blockIdx.x * static_cast<size_t>(kThreads) + threadIdx.x; i < per4;
             i += static_cast<size_t>(gridDim.x) * kThreads) {
            const float4 xv = x4[i];
            float4 r = make_float4(fmaf(a, xv.x, c), fmaf(a, xv.y, c), fmaf(a, xv.z, c), fmaf(a, xv.w, c));
            if (kHasY) {
                const float4 yv = y4[i];
                r.x = fmaf(b, clampf(yv.x, lim), r.x);
                r.y = fmaf(b, clampf(yv.y, lim), r.y);
                r.z = fmaf(b, clampf(yv.z, lim), r.z);
                r.w = fmaf(b, clampf(yv.w, lim), r.w);
            }
            o4[i] = r;
        }
    } else {
        for (size_t i = blockIdx.x * static_cast<size_t>(kThreads) + threadIdx.x; i < per;
             i += static_cast<size_t>(gridDim.x) * kThreads) {
            float r = fmaf(a, x[base + i], c);
            if (kHasY) r = fmaf(b, clampf(y[base + i], lim), r);
            out[base + i] = r;
        }
    }
}

// mode 0: out = clamp(x, lo, hi);  mode 1: out = g * [g * (x - clamp(x, lo, hi)) >= 0]
template <int kMode>
__global__ void __launch_bounds__(kThreads) clamp_grad_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                              float* __restrict__ out, size_t total, float lo, float hi) {
    for (size_t i = blockIdx.x * static_cast<size_t>(kThreads) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * kThreads) {
        const float xv = x[i];
        const float cl = (xv != xv) ? xv : fminf(fmaxf(xv, lo), hi);  // NaN propagates like torch.clamp (fminf / fmaxf drop it)
        if (kMode == 0) {
            out[i] = cl;
        } else {
            const float gv = g[i];
            out[i] = (gv * (xv - cl) >= 0.f) ? gv : 0.f;
        }
    }
}

}  // namespace
}  // namespace pcg

using namespace pcg;

extern "C" int pcg_affine2(const float* x, const float* y, const float* coef, float* out, int n_samples, long long per,
                           float clamp_limit, void* stream) {
    PCG_CHECK_ARG(x && coef && out, "pcg_affine2: null pointer");
    PCG_CHECK_ARG(n_samples > 0 && n_samples <= 65535 && per > 0, "pcg_affine2: bad shape n=%d per=%lld", n_samples, per);
    PCG_CHECK_ARG(y == nullptr || clamp_limit > 0.f, "pcg_affine2: clamp_limit must be positive (pass INFINITY for none)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool vec = (per % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                         reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
    const size_t work = vec ? per / 4 : per;
    const int bx = static_cast<int>(std::min<size_t>((work + kThreads - 1) / kThreads, 4 * static_cast<size_t>(sm_count())));
    const dim3 grid(bx > 0 ? bx : 1, n_samples);
    const size_t p = static_cast<size_t>(per);
    if (y != nullptr) {
        if (vec) affine2_kernel<true, true><<<grid, kThreads, 0, s>>>(x, y, coef, out, n_samples, p, clamp_limit);
        else affine2_kernel<true, false><<<grid, kThreads, 0, s>>>(x, y, coef, out, n_samples, p, clamp_limit);
    } else {
        if (vec) affine2_kernel<false, true><<<grid, kThreads, 0, s>>>(x, x, coef, out, n_samples, p, 0.f);
        else affine2_kernel<false, false><<<grid, kThreads, 0, s>>>(x, x, coef, out, n_samples, p, 0.f);
    }
    PCG_LAUNCH_CHECK("affine2_kernel");
    return 0;
}

extern "C" int pcg_clamp_with_grad(const float* x, const float* g, float* out, long long total, float lo, float hi,
                                   void* stream) {
    PCG_CHECK_ARG(x && out, "pcg_clamp_with_grad: null pointer");
    PCG_CHECK_ARG(total > 0 && lo <= hi, "pcg_clamp_with_grad: bad arguments total=%lld lo=%g hi=%g", total, lo, hi);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = static_cast<size_t>(total);
    const int blocks = static_cast<int>(std::min<size_t>((n + kThreads - 1) / kThreads, 8 * static_cast<size_t>(sm_count())));
    if (g == nullptr) clamp_grad_kernel<0><<<blocks, kThreads, 0, s>>>(x, nullptr, out, n, lo, hi);
    else clamp_grad_kernel<1><<<blocks, kThreads, 0, s>>>(x, g, out, n, lo, hi);
    PCG_LAUNCH_CHECK("clamp_grad_kernel");
    return 0;
}
