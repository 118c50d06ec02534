// Shared host-side plumbing for libpcg.so: error reporting across the C ABI, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/pcg.h"

namespace pcg {

// thread-local last-error message; no C++ exception ever crosses the ABI.
char* last_error_buf();
int set_error(int code, const char* fmt, ...);
void count_launch();
void reset_launch_count();
int launch_count();
int sm_count();

#define PCG_CHECK_ARG(cond, ...)                                    \
    do {                                                            \
        if (!(cond)) return ::pcg::set_error(-1, __VA_ARGS__);      \
    } while (0)

#define PCG_CUDA(expr)                                                                                 \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return ::pcg::set_error(static_cast<int>(_e), "%s failed: %s (%s:%d)", #expr,              \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);                       \
    } while (0)

// after a <<<>>> launch
#define PCG_LAUNCH_CHECK(name)                                                                         \
    do {                                                                                               \
        ::pcg::count_launch();                                                                         \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess)                                                                         \
            return ::pcg::set_error(static_cast<int>(_e), "launch of %s failed: %s", name,             \
                                    cudaGetErrorString(_e));                                           \
    } while (0)

// Optional per-kernel-family device timing (CUDA events on the launch stream), off by default.
// kinds: see PCG_PROF_* in include/pcg.h.  `work` is algorithmic FLOPs (tensor-bound kinds) or bytes (HBM-bound).
void profile_begin(int kind, double work, cudaStream_t stream);
void profile_end(cudaStream_t stream);
struct ProfileScope {
    cudaStream_t s;
    ProfileScope(int kind, double work, cudaStream_t stream) : s(stream) { profile_begin(kind, work, stream); }
    ~ProfileScope() { profile_end(s); }
};

// One-time per-DEVICE setup (cudaFuncSetAttribute is a per-device property: an encoder moved to a second GPU of the
// same process must configure its kernels there too), safe against concurrent first calls (autograd runs backward on
// its own thread).  `static PerDeviceOnce once; PCG_TRY_ONCE(once, cudaFuncSetAttribute(...));`
struct PerDeviceOnce {
    std::mutex mu;
    uint64_t done = 0;  // bit d: device ordinal d is configured (ordinals >= 64 are configured every time)
};
#define PCG_ONCE_PER_DEVICE(once, ...)                                          \
    do {                                                                        \
        int _dev = 0;                                                           \
        PCG_CUDA(cudaGetDevice(&_dev));                                         \
        std::lock_guard<std::mutex> _lock((once).mu);                           \
        if (_dev >= 64 || !(((once).done >> _dev) & 1ull)) {                    \
            __VA_ARGS__;                                                        \
            if (_dev < 64) (once).done |= 1ull << _dev;                         \
        }                                                                       \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

}  // namespace pcg
