// tcgen05 short-sequence attention for 128 <= T <= 272 (ViT-B/16: 197, ViT-L/14: 257): forward, dQ and dK/dV.
//
// One CTA = one (cutout, head).  Whole K and V (forward, dQ) or Q and dO (dK/dV) of the head sit in shared memory
// (TMA, 128-byte swizzle); the CTA walks 128-row tiles.  Every product is a tcgen05.mma with M = 128:
//     S = Q K^T           A = Q  (K-major)           B = K  (K-major, N = all keys, <= 272 = 256 + 16)
//     O = P V             A = P  (K-major, written by the softmax threads)   B = V (MN-major: rows = keys)
//     dP = dO V^T, dQ = dS K, and in the dK/dV kernel the transposed problem S^T = K Q^T, dP^T = V dO^T,
//     dV = P^T dO, dK = dS^T Q, so that the operand written by threads (P, dS, P^T, dS^T) is always K-major A.
// Accumulators live in tensor memory (S / dP: <= 272 columns, O / dQ / dV / dK: 64 columns each); 128 threads
// (thread = accumulator row = TMEM lane) do softmax / dS in registers with tcgen05.ld, one extra warp drives TMA
// and issues the MMAs.  The kernel is exp-bound (16 MUFU/clk/SM), not tensor-bound, which is why the phases are run
// back to back instead of being software pipelined.
// Rows beyond the last full 128-row tile (T = 257 -> one row) are handled by the mma.sync kernels in attention.cu.
//
// Replaces nn.MultiheadAttention's core (perceptor/models/ruclip/model.py:43-49) and its autograd.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "pcg_common.cuh"
#include "pcg_ptx.cuh"

namespace pcg {

int attn_fwd_legacy(const void* qkv, void* out, float* lse, int n, int T, int heads, int q_begin, cudaStream_t s);
int attn_delta(const void* out, const void* d_out, float* delta, int n, int T, int heads, cudaStream_t s);
int attn_bwd_legacy(const void* qkv, const void* d_out, const float* lse, const float* delta, void* d_qkv, int n, int T,
                    int heads, int begin, cudaStream_t s);

namespace {

using bf16 = __nv_bfloat16;
constexpr int kHd = 64;
constexpr int kRowsMax = 320;                 // 5 blocks of 64 rows: holds up to 272 (+ padding) keys / queries
constexpr int kBlkBytes = 128 * 128;          // one [128 rows x 64 bf16] swizzled block = 16 KB
constexpr int kLongBytes = kRowsMax * 128;    // a whole-head operand (K, V, Q or dO): 40 KB
constexpr int kPBytes = 5 * kBlkBytes;        // P / dS: [128 x 320] bf16 as 5 K-major blocks = 80 KB
constexpr int kThreads = 160;                 // 4 compute warps (TMEM lane quarters) + 1 control warp
constexpr float kLog2e = 1.4426950408889634f;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColAcc0 = 320;            // first 64-column accumulator (O / dQ / dV)
constexpr uint32_t kColAcc1 = 384;            // second 64-column accumulator (dK)

struct TcParams {
    int T, heads, nk;       // nk = T rounded up to 16 (MMA N / K extent over keys or queries)
    int n_tiles;            // 128-row tiles handled here
    bf16* out;              // fwd: [n*T, D]
    float* lse;             // [n, heads, T]
    const float* delta;     // [n, heads, T]
    bf16* d_qkv;            // [n*T, 3D]
};

// byte offset of the 16-byte chunk `chunk` (8 bf16) of row `row` inside a K-major operand made of [128 x 64] blocks
__device__ __forceinline__ uint32_t p_offset(int row, int chunk) {
    return static_cast<uint32_t>((chunk >> 3) * kBlkBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

// S-type product: D[128, nk] = A[128 x 64] (K-major) * B[nk x 64]^T (K-major), nk <= 272 split as 256 + rest
__device__ __forceinline__ void mma_rows_x_long(uint32_t tmem_d, const uint8_t* a_tile, const uint8_t* b_long, int nk) {
    const int n0 = nk < 256 ? nk : 256;
    const int n1 = nk - n0;
    const uint32_t idesc0 = umma_idesc_bf16(128, n0);
    const uint32_t idesc1 = umma_idesc_bf16(128, n1 > 0 ? n1 : 16);
    const uint64_t da = umma_smem_desc_sw128(smem_u32(a_tile));
    const uint64_t db0 = umma_smem_desc_sw128(smem_u32(b_long));
    const uint64_t db1 = umma_smem_desc_sw128(smem_u32(b_long + 256 * 128));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        umma_f16(tmem_d, da + 2 * k, db0 + 2 * k, idesc0, k != 0);
        if (n1 > 0) umma_f16(tmem_d + 256, da + 2 * k, db1 + 2 * k, idesc1, k != 0);
    }
}

// PV-type product: D[128, 64] = A[128 x nk] (K-major blocks written by threads) * B[nk x 64] (MN-major: rows = k)
__device__ __forceinline__ void mma_p_x_rows(uint32_t tmem_d, const uint8_t* p_buf, const uint8_t* b_long, int nk) {
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
    const int ksteps = nk >> 4;
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = umma_smem_desc_sw128(smem_u32(p_buf + (ks >> 2) * kBlkBytes + (ks & 3) * 32));
        const uint64_t db = umma_smem_desc_sw128(smem_u32(b_long + ks * 2048));
        umma_f16(tmem_d, da, db, idesc, ks != 0);
    }
}

struct Shared {
    uint8_t* tile_a;   // [128 x 64] Q (fwd, dq) / K (dkdv)
    uint8_t* tile_b;   // [128 x 64] dO (dq) / V (dkdv)
    uint8_t* long_a;   // whole-head K (fwd, dq) / Q (dkdv)
    uint8_t* long_b;   // whole-head V (fwd, dq) / dO (dkdv)
    uint8_t* p_buf;    // P / dS
    float* vec_a;      // [320] lse (dkdv)
    float* vec_b;      // [320] delta (dkdv)
    uint64_t* bars;    // [0] long operands, [1] tile operands, [2] mma A, [3] mma B, [4] mma C
    uint32_t* tmem_slot;
};
constexpr int kSmemBytes = 2 * kBlkBytes + 2 * kLongBytes + kPBytes + 2 * kRowsMax * 4 + 64 + 1024;

__device__ __forceinline__ Shared carve(uint8_t* raw) {
    const uint32_t addr = smem_u32(raw);
    uint8_t* base = raw + (((addr + 1023u) & ~1023u) - addr);
    Shared s;
    s.tile_a = base;
    s.tile_b = base + kBlkBytes;
    s.long_a = base + 2 * kBlkBytes;
    s.long_b = s.long_a + kLongBytes;
    s.p_buf = s.long_b + kLongBytes;
    s.vec_a = reinterpret_cast<float*>(s.p_buf + kPBytes);
    s.vec_b = s.vec_a + kRowsMax;
    s.bars = reinterpret_cast<uint64_t*>(s.vec_b + kRowsMax);
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.bars + 6);
    return s;
}

// load rows [0, nk) of a whole-head operand: full 128-row boxes, then one `tail`-row box (tail = nk % 128)
__device__ __forceinline__ void tma_long(const CUtensorMap* map128, const CUtensorMap* map_tail, uint64_t* bar,
                                         uint8_t* dst, int col, int n, int nk) {
    const int full = nk >> 7;
    for (int i = 0; i < full; ++i) tma_load_3d(map128, bar, dst + i * kBlkBytes, col, i * 128, n, kEvictNormal);
    if (nk & 127) tma_load_3d(map_tail, bar, dst + full * kBlkBytes, col, full * 128, n, kEvictNormal);
}

__device__ __forceinline__ void setup(const Shared& sm, int warp, int lane) {
    if (warp == 4) {
        if (lane == 0) {
            for (int i = 0; i < 5; ++i) mbar_init(&sm.bars[i], 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(sm.tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

// write one accumulator row (64 fp32 in TMEM columns [col, col+64)) * scale as bf16 into row `r` of a swizzled
// [128 x 64] staging tile; the caller then copies whole rows out with coalesced 16-byte stores.
__device__ __forceinline__ void acc_row_to_tile(uint32_t taddr, float scale, uint8_t* tile, int r) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 q = make_uint4(
                pack_bf16(__uint_as_float(v[8 * j]) * scale, __uint_as_float(v[8 * j + 1]) * scale),
                pack_bf16(__uint_as_float(v[8 * j + 2]) * scale, __uint_as_float(v[8 * j + 3]) * scale),
                pack_bf16(__uint_as_float(v[8 * j + 4]) * scale, __uint_as_float(v[8 * j + 5]) * scale),
                pack_bf16(__uint_as_float(v[8 * j + 6]) * scale, __uint_as_float(v[8 * j + 7]) * scale));
            *reinterpret_cast<uint4*>(tile + r * 128 + (((c * 4 + j) ^ (r & 7)) << 4)) = q;
        }
    }
}
// each warp copies its own 32 rows of the staging tile to global rows t0 + row (row stride ld elements)
__device__ __forceinline__ void tile_rows_to_global(const uint8_t* tile, int warp, int lane, bf16* gbase, size_t ld, int t0,
                                                    int T) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = warp * 32 + i * 4 + (lane >> 3), ch = lane & 7;
        if (t0 + row < T)
            *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(t0 + row) * ld + ch * 8) =
                *reinterpret_cast<const uint4*>(tile + row * 128 + ((ch ^ (row & 7)) << 4));
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------
// forward: O = softmax(Q K^T) V, lse
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map_tail,
                   const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const Shared sm = carve(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, n = blockIdx.y;
    const int D = p.heads * kHd;
    setup(sm, warp, lane);
    const uint32_t tmem = *sm.tmem_slot;
    const int nk = p.nk;

    if (warp == 4 && lane == 0) {
        mbar_arrive_expect_tx(&sm.bars[0], 2 * nk * 128);
        tma_long(&map128, &map_tail, &sm.bars[0], sm.long_a, D + h * kHd, n, nk);      // K
        tma_long(&map128, &map_tail, &sm.bars[0], sm.long_b, 2 * D + h * kHd, n, nk);  // V
    }
    for (int qt = 0; qt < p.n_tiles; ++qt) {
        const uint32_t ph = qt & 1;
        const int q0 = qt * 128;
        if (warp == 4) {
            if (lane == 0) {
                mbar_arrive_expect_tx(&sm.bars[1], kBlkBytes);
                tma_load_3d(&map128, &sm.bars[1], sm.tile_a, h * kHd, q0, n, kEvictFirst);  // Q tile
                if (qt == 0) mbar_wait(&sm.bars[0], 0);
                mbar_wait(&sm.bars[1], ph);
                tc_fence_after();
                mma_rows_x_long(tmem, sm.tile_a, sm.long_a, nk);  // S = Q K^T
                umma_commit(&sm.bars[2]);
            }
        } else {
            const int r = warp * 32 + lane;  // accumulator row == TMEM lane
            const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
            mbar_wait(&sm.bars[2], ph);
            tc_fence_after();
            // pass 1: row max over the valid keys
            float mx = -INFINITY;
            for (int c = 0; c < nk; c += 32) {
                if (c + 32 <= nk) {
                    uint32_t v[32];
                    tmem_ld_32x32(trow + c, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c + j < p.T) ? __uint_as_float(v[j]) : -INFINITY);
                } else {
                    uint32_t v[16];
                    tmem_ld_32x16(trow + c, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, (c + j < p.T) ? __uint_as_float(v[j]) : -INFINITY);
                }
            }
            // pass 2: P = exp(S - max) as bf16 into the K-major P buffer, row sum
            const float mb = mx * kLog2e;
            float sum = 0.f;
            for (int c = 0; c < nk; c += 32) {
                uint32_t v[32];
                const bool full = c + 32 <= nk;
                if (full) {
                    tmem_ld_32x32(trow + c, v);
                } else {
                    uint32_t w[16];
                    tmem_ld_32x16(trow + c, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = w[j], v[16 + j] = 0u;
                }
                tmem_wait_ld();
                float e[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    e[j] = (c + j < p.T) ? exp2f(fmaf(__uint_as_float(v[j]), kLog2e, -mb)) : 0.f;
                    sum += e[j];
                }
                const int nchunk = full ? 4 : 2;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nchunk)
                        *reinterpret_cast<uint4*>(sm.p_buf + p_offset(r, (c >> 3) + j)) =
                            make_uint4(pack_bf16(e[8 * j], e[8 * j + 1]), pack_bf16(e[8 * j + 2], e[8 * j + 3]),
                                       pack_bf16(e[8 * j + 4], e[8 * j + 5]), pack_bf16(e[8 * j + 6], e[8 * j + 7]));
            }
            if (q0 + r < p.T) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + q0 + r] = mx + logf(sum);
            // keep 1/sum for the O epilogue in a register across the barrier
            asm volatile("" ::"f"(sum));
            fence_proxy_async();  // P was written through the generic proxy, the MMA reads it through the async proxy
            tc_fence_before();
            __syncthreads();
            // (control warp issues O = P V here)
            mbar_wait(&sm.bars[3], ph);
            tc_fence_after();
            acc_row_to_tile(trow + kColAcc0, 1.0f / sum, sm.tile_a, r);
            tile_rows_to_global(sm.tile_a, warp, lane, p.out + static_cast<size_t>(n) * p.T * D + h * kHd, D, q0, p.T);
        }
        if (warp == 4) {
            tc_fence_before();
            __syncthreads();  // P complete
            if (lane == 0) {
                tc_fence_after();
                mma_p_x_rows(tmem + kColAcc0, sm.p_buf, sm.long_b, nk);  // O = P V
                umma_commit(&sm.bars[3]);
            }
        }
        tc_fence_before();
        __syncthreads();  // O drained, Q tile buffer reusable
        tc_fence_after();
    }
    if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// backward, dQ:  S = Q K^T, P = exp(S - lse), dP = dO V^T, dS = P (dP - delta), dQ = dS K
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
attn_dq_tc_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map_tail,
                  const __grid_constant__ CUtensorMap map_do128, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const Shared sm = carve(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, n = blockIdx.y;
    const int D = p.heads * kHd;
    setup(sm, warp, lane);
    const uint32_t tmem = *sm.tmem_slot;
    const int nk = p.nk;

    if (warp == 4 && lane == 0) {
        mbar_arrive_expect_tx(&sm.bars[0], 2 * nk * 128);
        tma_long(&map128, &map_tail, &sm.bars[0], sm.long_a, D + h * kHd, n, nk);      // K
        tma_long(&map128, &map_tail, &sm.bars[0], sm.long_b, 2 * D + h * kHd, n, nk);  // V
    }
    for (int qt = 0; qt < p.n_tiles; ++qt) {
        const uint32_t ph = qt & 1;
        const int q0 = qt * 128;
        if (warp == 4) {
            if (lane == 0) {
                mbar_arrive_expect_tx(&sm.bars[1], 2 * kBlkBytes);
                tma_load_3d(&map128, &sm.bars[1], sm.tile_a, h * kHd, q0, n, kEvictFirst);     // Q tile
                tma_load_3d(&map_do128, &sm.bars[1], sm.tile_b, h * kHd, q0, n, kEvictFirst);  // dO tile
                if (qt == 0) mbar_wait(&sm.bars[0], 0);
                mbar_wait(&sm.bars[1], ph);
                tc_fence_after();
                mma_rows_x_long(tmem, sm.tile_a, sm.long_a, nk);  // S = Q K^T
                umma_commit(&sm.bars[2]);
            }
            tc_fence_before();
            __syncthreads();  // (1) P written, S consumed
            if (lane == 0) {
                tc_fence_after();
                mma_rows_x_long(tmem, sm.tile_b, sm.long_b, nk);  // dP = dO V^T (reuses the S columns)
                umma_commit(&sm.bars[3]);
            }
            tc_fence_before();
            __syncthreads();  // (2) dS written
            if (lane == 0) {
                tc_fence_after();
                mma_p_x_rows(tmem + kColAcc0, sm.p_buf, sm.long_a, nk);  // dQ = dS K
                umma_commit(&sm.bars[4]);
            }
        } else {
            const int r = warp * 32 + lane;
            const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
            const bool row_ok = q0 + r < p.T;
            const size_t vidx = (static_cast<size_t>(n) * p.heads + h) * p.T + q0 + r;
            const float lse_b = row_ok ? p.lse[vidx] * kLog2e : INFINITY;
            const float dlt = row_ok ? p.delta[vidx] : 0.f;
            mbar_wait(&sm.bars[2], ph);
            tc_fence_after();
            // P = exp(S - lse)
            for (int c = 0; c < nk; c += 32) {
                uint32_t v[32];
                const bool full = c + 32 <= nk;
                if (full) {
                    tmem_ld_32x32(trow + c, v);
                } else {
                    uint32_t w[16];
                    tmem_ld_32x16(trow + c, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = w[j], v[16 + j] = 0u;
                }
                tmem_wait_ld();
                float e[32];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    e[j] = (c + j < p.T) ? exp2f(fmaf(__uint_as_float(v[j]), kLog2e, -lse_b)) : 0.f;
                const int nchunk = full ? 4 : 2;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nchunk)
                        *reinterpret_cast<uint4*>(sm.p_buf + p_offset(r, (c >> 3) + j)) =
                            make_uint4(pack_bf16(e[8 * j], e[8 * j + 1]), pack_bf16(e[8 * j + 2], e[8 * j + 3]),
                                       pack_bf16(e[8 * j + 4], e[8 * j + 5]), pack_bf16(e[8 * j + 6], e[8 * j + 7]));
            }
            tc_fence_before();
            __syncthreads();  // (1)
            mbar_wait(&sm.bars[3], ph);
            tc_fence_after();
            // dS = P * (dP - delta), in place over P
            for (int c = 0; c < nk; c += 32) {
                uint32_t v[32];
                const bool full = c + 32 <= nk;
                if (full) {
                    tmem_ld_32x32(trow + c, v);
                } else {
                    uint32_t w[16];
                    tmem_ld_32x16(trow + c, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = w[j], v[16 + j] = 0u;
                }
                tmem_wait_ld();
                const int nchunk = full ? 4 : 2;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nchunk) {
                        uint4* slot = reinterpret_cast<uint4*>(sm.p_buf + p_offset(r, (c >> 3) + j));
                        const uint4 pv = *slot;
                        const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
                        uint32_t o[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const __nv_bfloat162 p2 = *reinterpret_cast<const __nv_bfloat162*>(&pw[q]);
                            o[q] = pack_bf16(__low2float(p2) * (__uint_as_float(v[8 * j + 2 * q]) - dlt),
                                             __high2float(p2) * (__uint_as_float(v[8 * j + 2 * q + 1]) - dlt));
                        }
                        *slot = make_uint4(o[0], o[1], o[2], o[3]);
                    }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();  // (2)
            mbar_wait(&sm.bars[4], ph);
            tc_fence_after();
            acc_row_to_tile(trow + kColAcc0, 1.0f, sm.tile_a, r);
            tile_rows_to_global(sm.tile_a, warp, lane, p.d_qkv + static_cast<size_t>(n) * p.T * 3 * D + h * kHd, 3 * D, q0,
                                p.T);
        }
        tc_fence_before();
        __syncthreads();  // dQ drained; tile buffers and P reusable
        tc_fence_after();
    }
    if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// backward, dK / dV on the transposed problem (rows = keys):
//   S^T = K Q^T, P^T = exp(S^T - lse[q]), dP^T = V dO^T, dS^T = P^T (dP^T - delta[q]), dV = P^T dO, dK = dS^T Q
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
attn_dkdv_tc_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map_tail,
                    const __grid_constant__ CUtensorMap map_do128, const __grid_constant__ CUtensorMap map_do_tail,
                    const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const Shared sm = carve(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, n = blockIdx.y;
    const int D = p.heads * kHd;
    setup(sm, warp, lane);
    const uint32_t tmem = *sm.tmem_slot;
    const int nk = p.nk;  // here: padded number of QUERIES (MMA N of S^T, K of dV / dK)

    if (warp == 4 && lane == 0) {
        mbar_arrive_expect_tx(&sm.bars[0], 2 * nk * 128);
        tma_long(&map128, &map_tail, &sm.bars[0], sm.long_a, h * kHd, n, nk);        // Q (all queries)
        tma_long(&map_do128, &map_do_tail, &sm.bars[0], sm.long_b, h * kHd, n, nk);  // dO
    }
    {
        const size_t vbase = (static_cast<size_t>(n) * p.heads + h) * p.T;
        for (int i = threadIdx.x; i < kRowsMax; i += kThreads) {
            sm.vec_a[i] = (i < p.T) ? p.lse[vbase + i] * kLog2e : INFINITY;
            sm.vec_b[i] = (i < p.T) ? p.delta[vbase + i] : 0.f;
        }
    }
    __syncthreads();
    for (int kt = 0; kt < p.n_tiles; ++kt) {
        const uint32_t ph = kt & 1;
        const int k0 = kt * 128;
        if (warp == 4) {
            if (lane == 0) {
                mbar_arrive_expect_tx(&sm.bars[1], 2 * kBlkBytes);
                tma_load_3d(&map128, &sm.bars[1], sm.tile_a, D + h * kHd, k0, n, kEvictFirst);      // K tile
                tma_load_3d(&map128, &sm.bars[1], sm.tile_b, 2 * D + h * kHd, k0, n, kEvictFirst);  // V tile
                if (kt == 0) mbar_wait(&sm.bars[0], 0);
                mbar_wait(&sm.bars[1], ph);
                tc_fence_after();
                mma_rows_x_long(tmem, sm.tile_a, sm.long_a, nk);  // S^T = K Q^T
                umma_commit(&sm.bars[2]);
            }
            tc_fence_before();
            __syncthreads();  // (1) P^T written, S^T consumed
            if (lane == 0) {
                tc_fence_after();
                mma_rows_x_long(tmem, sm.tile_b, sm.long_b, nk);               // dP^T = V dO^T
                mma_p_x_rows(tmem + kColAcc0, sm.p_buf, sm.long_b, nk);        // dV = P^T dO
                umma_commit(&sm.bars[3]);
            }
            tc_fence_before();
            __syncthreads();  // (2) dS^T written (after both MMAs above completed)
            if (lane == 0) {
                tc_fence_after();
                mma_p_x_rows(tmem + kColAcc1, sm.p_buf, sm.long_a, nk);  // dK = dS^T Q
                umma_commit(&sm.bars[4]);
            }
        } else {
            const int r = warp * 32 + lane;  // key row within the tile
            const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
            mbar_wait(&sm.bars[2], ph);
            tc_fence_after();
            for (int c = 0; c < nk; c += 32) {
                uint32_t v[32];
                const bool full = c + 32 <= nk;
                if (full) {
                    tmem_ld_32x32(trow + c, v);
                } else {
                    uint32_t w[16];
                    tmem_ld_32x16(trow + c, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = w[j], v[16 + j] = 0u;
                }
                tmem_wait_ld();
                float e[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) e[j] = exp2f(fmaf(__uint_as_float(v[j]), kLog2e, -sm.vec_a[c + j]));
                const int nchunk = full ? 4 : 2;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nchunk)
                        *reinterpret_cast<uint4*>(sm.p_buf + p_offset(r, (c >> 3) + j)) =
                            make_uint4(pack_bf16(e[8 * j], e[8 * j + 1]), pack_bf16(e[8 * j + 2], e[8 * j + 3]),
                                       pack_bf16(e[8 * j + 4], e[8 * j + 5]), pack_bf16(e[8 * j + 6], e[8 * j + 7]));
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();  // (1)
            mbar_wait(&sm.bars[3], ph);  // dP^T ready and dV finished reading P^T
            tc_fence_after();
            for (int c = 0; c < nk; c += 32) {
                uint32_t v[32];
                const bool full = c + 32 <= nk;
                if (full) {
                    tmem_ld_32x32(trow + c, v);
                } else {
                    uint32_t w[16];
                    tmem_ld_32x16(trow + c, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = w[j], v[16 + j] = 0u;
                }
                tmem_wait_ld();
                const int nchunk = full ? 4 : 2;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nchunk) {
                        uint4* slot = reinterpret_cast<uint4*>(sm.p_buf + p_offset(r, (c >> 3) + j));
                        const uint4 pv = *slot;
                        const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
                        uint32_t o[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int col = c + 8 * j + 2 * q;
                            const __nv_bfloat162 p2 = *reinterpret_cast<const __nv_bfloat162*>(&pw[q]);
                            o[q] = pack_bf16(__low2float(p2) * (__uint_as_float(v[8 * j + 2 * q]) - sm.vec_b[col]),
                                             __high2float(p2) * (__uint_as_float(v[8 * j + 2 * q + 1]) - sm.vec_b[col + 1]));
                        }
                        *slot = make_uint4(o[0], o[1], o[2], o[3]);
                    }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();  // (2)
            // dV is complete (bars[3]); write it out while dK is being computed
            bf16* gd = p.d_qkv + static_cast<size_t>(n) * p.T * 3 * D + h * kHd;
            acc_row_to_tile(trow + kColAcc0, 1.0f, sm.tile_b, r);
            tile_rows_to_global(sm.tile_b, warp, lane, gd + 2 * D, 3 * D, k0, p.T);
            mbar_wait(&sm.bars[4], ph);
            tc_fence_after();
            acc_row_to_tile(trow + kColAcc1, 1.0f, sm.tile_a, r);
            tile_rows_to_global(sm.tile_a, warp, lane, gd + D, 3 * D, k0, p.T);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
using EncodeFn = PFN_cuTensorMapEncodeTiled_v12000;
EncodeFn encode_fn() {
    static EncodeFn fn = []() -> EncodeFn {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeFn>(ptr);
    }();
    return fn;
}

// [n][T][cols] bf16 view of a row-major [n*T, cols] matrix; box = {64 columns, box_rows, 1}
int make_map3(CUtensorMap* out, const void* ptr, int n, int T, int cols, int box_rows) {
    struct Key {
        const void* p;
        int n, T, cols, box;
        bool operator==(const Key& o) const { return p == o.p && n == o.n && T == o.T && cols == o.cols && box == o.box; }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            size_t h = reinterpret_cast<size_t>(k.p);
            for (int v : {k.n, k.T, k.cols, k.box}) h = h * 1000003u ^ static_cast<size_t>(v);
            return h;
        }
    };
    static std::mutex mu;
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    const Key key{ptr, n, T, cols, box_rows};
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeFn encode = encode_fn();
    if (encode == nullptr) return set_error(-2, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(n)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(cols) * 2 * T};
    const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(-3, "cuTensorMapEncodeTiled(3d) failed: CUresult %d", static_cast<int>(r));
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

struct TcPlan {
    bool use_tc;
    int nk, n_tiles, legacy_begin;
};
TcPlan plan_tc(int T) {
    TcPlan pl{false, 0, 0, 0};
    if (T < 128 || T > 272) return pl;
    pl.use_tc = true;
    pl.nk = (T + 15) / 16 * 16;
    const int full = T / 128, tail = T - full * 128;
    if (tail > 64) {  // a partially filled tcgen05 tile beats two mma.sync tiles
        pl.n_tiles = full + 1;
        pl.legacy_begin = T;
    } else {
        pl.n_tiles = full;
        pl.legacy_begin = full * 128;
    }
    return pl;
}

template <typename K>
int set_smem(K kernel) {
    PCG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return 0;
}

// The tcgen05 path is validated but, run phase by phase with one CTA per SM, it is still slower than the mma.sync
// kernels on B200 (attention section): it is opt-in (PCG_ATTN_TC=1 or pcg_attn_set_legacy(0)).
bool g_disable_tc = []() {
    const char* e = getenv("PCG_ATTN_TC");
    return !(e != nullptr && e[0] == '1');
}();

}  // namespace
}  // namespace pcg

using namespace pcg;

extern "C" int pcg_attn_set_legacy(int on) {  // test hook: force the mma.sync kernels for every row
    g_disable_tc = on != 0;
    return 0;
}

extern "C" int pcg_attn_fwd(const void* qkv, void* out, float* lse, int n, int T, int heads, void* stream) {
    PCG_CHECK_ARG(qkv && out && lse, "pcg_attn_fwd: null pointer");
    PCG_CHECK_ARG(n > 0 && T > 0 && heads > 0, "pcg_attn_fwd: bad shape n=%d T=%d heads=%d", n, T, heads);
    PCG_CHECK_ARG(n <= 65535 && heads <= 65535, "pcg_attn_fwd: n and heads must be <= 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_FWD, 4.0 * T * T * kHd * heads * n, s);
    const TcPlan pl = g_disable_tc ? TcPlan{false, 0, 0, 0} : plan_tc(T);
    if (!pl.use_tc) return attn_fwd_legacy(qkv, out, lse, n, T, heads, 0, s);
    const int D = heads * kHd;
    CUtensorMap m128, mtail;
    if (int rc = make_map3(&m128, qkv, n, T, 3 * D, 128)) return rc;
    if (int rc = make_map3(&mtail, qkv, n, T, 3 * D, (pl.nk & 127) ? (pl.nk & 127) : 128)) return rc;
    static bool configured = false;
    if (!configured) {
        if (int rc = set_smem(attn_fwd_tc_kernel)) return rc;
        configured = true;
    }
    TcParams p{T, heads, pl.nk, pl.n_tiles, static_cast<bf16*>(out), lse, nullptr, nullptr};
    attn_fwd_tc_kernel<<<dim3(heads, n), kThreads, kSmemBytes, s>>>(m128, mtail, p);
    PCG_LAUNCH_CHECK("attn_fwd_tc_kernel");
    return attn_fwd_legacy(qkv, out, lse, n, T, heads, pl.legacy_begin, s);
}

extern "C" int pcg_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                            void* d_qkv, int n, int T, int heads, void* stream) {
    PCG_CHECK_ARG(qkv && out && d_out && lse && delta_ws && d_qkv, "pcg_attn_bwd: null pointer");
    PCG_CHECK_ARG(n > 0 && T > 0 && heads > 0, "pcg_attn_bwd: bad shape n=%d T=%d heads=%d", n, T, heads);
    PCG_CHECK_ARG(n <= 65535 && heads <= 65535, "pcg_attn_bwd: n and heads must be <= 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_BWD, 8.0 * T * T * kHd * heads * n, s);
    if (int rc = attn_delta(out, d_out, delta_ws, n, T, heads, s)) return rc;
    const TcPlan pl = g_disable_tc ? TcPlan{false, 0, 0, 0} : plan_tc(T);
    if (!pl.use_tc) return attn_bwd_legacy(qkv, d_out, lse, delta_ws, d_qkv, n, T, heads, 0, s);
    const int D = heads * kHd;
    const int tail = (pl.nk & 127) ? (pl.nk & 127) : 128;
    CUtensorMap m128, mtail, mdo128, mdotail;
    if (int rc = make_map3(&m128, qkv, n, T, 3 * D, 128)) return rc;
    if (int rc = make_map3(&mtail, qkv, n, T, 3 * D, tail)) return rc;
    if (int rc = make_map3(&mdo128, d_out, n, T, D, 128)) return rc;
    if (int rc = make_map3(&mdotail, d_out, n, T, D, tail)) return rc;
    static bool configured = false;
    if (!configured) {
        if (int rc = set_smem(attn_dq_tc_kernel)) return rc;
        if (int rc = set_smem(attn_dkdv_tc_kernel)) return rc;
        configured = true;
    }
    TcParams p{T, heads, pl.nk, pl.n_tiles, nullptr, const_cast<float*>(lse), delta_ws, static_cast<bf16*>(d_qkv)};
    attn_dkdv_tc_kernel<<<dim3(heads, n), kThreads, kSmemBytes, s>>>(m128, mtail, mdo128, mdotail, p);
    PCG_LAUNCH_CHECK("attn_dkdv_tc_kernel");
    attn_dq_tc_kernel<<<dim3(heads, n), kThreads, kSmemBytes, s>>>(m128, mtail, mdo128, p);
    PCG_LAUNCH_CHECK("attn_dq_tc_kernel");
    return attn_bwd_legacy(qkv, d_out, lse, delta_ws, d_qkv, n, T, heads, pl.legacy_begin, s);
}
