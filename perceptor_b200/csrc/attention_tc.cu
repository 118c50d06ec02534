// tcgen05 attention for ViT sequences, head dim 64.  Two kernel pairs:
//   * 66 <= T <= 257 (ViT-B/16: 197, ViT-L/14: 257): forward with all of S in tensor memory and a single fused
//     backward (dQ, dK, dV in one pass, S and dP computed once), both built on the decomposition below;
//   * 257 < T <= 1152 (ViT-L/14 @336: 577): a streaming forward (online softmax over 128-key chunks) and a backward
//     with one CTA per key tile whose partial dQ is reduced with fp32 vector reductions; see the sections further down.
//
// "256 + 1" decomposition.  A ViT sequence is a power-of-two patch grid plus the class token, so T - 1 tokens
// tile exactly into 128-row tensor-core tiles and the one left-over token (the "edge" token x = T - 1, both as a
// query and as a key) is a handful of 64-long dot products on the CUDA cores:
//   * tensor cores: tokens [0, T-1) against tokens [0, T-1); every accumulator lives in tensor memory,
//   * CUDA cores:   row x and column x of the score matrix (matrix-vector products against operands that already
//                   sit in shared memory), folded into the epilogues.
// That keeps S at <= 256 TMEM columns, so two forward CTAs share an SM (the softmax of one overlaps the MMAs of
// the other), and the fused backward fits its six accumulators in the 512 columns of one SM.
//
// forward  (CTA = cutout, head, 128-query tile; 2 CTAs / SM):
//   S = Q K^T (SS MMA, N = keys) -> softmax in registers (thread = row = TMEM lane) -> P written back to TMEM as
//   packed bf16 over the dead S columns -> O = P V with A = P read from TMEM (TS MMA) -> epilogue adds the edge
//   key's p_x v_x, scales by 1 / sum, stores bf16.  An extra warp computes the edge query row.
// backward (CTA = cutout, head; 1 CTA / SM), transposed problem, rows = keys, 128 x 128 blocks (key tile j, query
//   block i):  S^T = K_j Q_i^T, dP^T = V_j dO_i^T  ->  P^T = exp2(S^T - lse_i), dS^T = P^T (dP^T - delta_i) written
//   as bf16 to shared memory  ->  dV_j += P^T dO_i, dK_j += dS^T Q_i (K-major A), dQ_i += dS K_j (the same dS^T
//   buffer read as an MN-major A).  TMEM: S^T 128 | dP^T 128 | dV 64 | dK 64 | dQ_0 64 | dQ_1 64 = 512 columns.
// Measured rates that shaped this (tools/ubench.cu, B200): tcgen05.ld 1.2 KB/clk/SM with 4 warps (not a limit),
// ex2 16/clk/SM (the softmax bound), one tcgen05.mma costs >= 96 clk whatever N is (so the N = 64 products are
// issue-bound at a third of peak), N = 256 runs at 75 % (A from smem) / 92 % (A from TMEM) of peak.
//
// Replaces nn.MultiheadAttention's core (perceptor/models/ruclip/model.py:43-49) and its autograd.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <algorithm>
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
constexpr int kBlkBytes = 128 * 128;  // one [128 rows x 64 bf16] 128B-swizzled block = 16 KB
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// byte offset of the 16-byte chunk `chunk` (0..7) of row `row` in a swizzled [rows x 64] bf16 operand
__device__ __forceinline__ uint32_t row_chunk(int row, int chunk) {
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// dot product of a swizzled smem row (64 bf16) with a float[64] vector in shared memory (4 independent chains)
// (not inlined, like edge_gemv below: the 256 + 1 kernels are instruction-cache bound -- ncu: 17 % of stalls are
// no_instructions -- and these run off the critical path; inlining them made the forward 9 % slower)
__device__ __noinline__ float row_dot(const uint8_t* mat, int row, const float* vec) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 m = *reinterpret_cast<const uint4*>(mat + row_chunk(row, c));
        const float4 a = *reinterpret_cast<const float4*>(vec + c * 8);
        const float4 b = *reinterpret_cast<const float4*>(vec + c * 8 + 4);
        a0 = fmaf(bf_lo(m.x), a.x, a0);
        a1 = fmaf(bf_hi(m.x), a.y, a1);
        a2 = fmaf(bf_lo(m.y), a.z, a2);
        a3 = fmaf(bf_hi(m.y), a.w, a3);
        a0 = fmaf(bf_lo(m.z), b.x, a0);
        a1 = fmaf(bf_hi(m.z), b.y, a1);
        a2 = fmaf(bf_lo(m.w), b.z, a2);
        a3 = fmaf(bf_hi(m.w), b.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one 64-long row of a [rows, ld] bf16 matrix -> float[64] in shared memory (one warp, 2 elements per lane)
__device__ __forceinline__ void load_row_f32(float* dst, const bf16* src, int lane) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(src + 2 * lane);
    dst[2 * lane] = bf_lo(u);
    dst[2 * lane + 1] = bf_hi(u);
}

// out[0..63] = (sum_i coef[i] * mat[i][.] + corner * xrow[.]) * scale as bf16, one warp.  mat = swizzled smem rows,
// coef / xrow = float vectors in shared memory.  Lane = (row group lane >> 3, 16-byte chunk lane & 7): the four row
// groups each walk a quarter of the rows with 8 independent accumulators and are summed by two shuffles at the end.
__device__ __noinline__ void edge_gemv(const float* coef, const uint8_t* mat, int rows, float corner,
                                          const float* xrow, bf16* gdst, float scale, int lane) {
    const int kg = lane >> 3, ch = lane & 7;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
    for (int i = kg; i < rows; i += 4) {
        const float c = coef[i];
        const uint4 m = *reinterpret_cast<const uint4*>(mat + row_chunk(i, ch));
        acc[0] = fmaf(c, bf_lo(m.x), acc[0]);
        acc[1] = fmaf(c, bf_hi(m.x), acc[1]);
        acc[2] = fmaf(c, bf_lo(m.y), acc[2]);
        acc[3] = fmaf(c, bf_hi(m.y), acc[3]);
        acc[4] = fmaf(c, bf_lo(m.z), acc[4]);
        acc[5] = fmaf(c, bf_hi(m.z), acc[5]);
        acc[6] = fmaf(c, bf_lo(m.w), acc[6]);
        acc[7] = fmaf(c, bf_hi(m.w), acc[7]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    }
    if (kg == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(corner, xrow[ch * 8 + k], acc[k]) * scale;
        *reinterpret_cast<uint4*>(gdst + ch * 8) = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]),
                                                              pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
    }
}

// D[128, n] = A[128 x 64] * B[n x 64]^T, both K-major swizzled tiles (n <= 256: B rows run on across 16 KB blocks)
__device__ __forceinline__ void mma_tile_x_rows(uint32_t tmem_d, const uint8_t* a_tile, const uint8_t* b_rows, int n) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint64_t da = umma_smem_desc_sw128(smem_u32(a_tile));
    const uint64_t db = umma_smem_desc_sw128(smem_u32(b_rows));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, k != 0);
}

// D[128, 64] (+)= A[128 x 16*ksteps] (K-major 64-column blocks written by threads) * B[16*ksteps x 64] (rows = k)
__device__ __forceinline__ void mma_blocks_x_cols(uint32_t tmem_d, const uint8_t* a_blocks, const uint8_t* b_rows,
                                                  int ksteps, bool accumulate) {
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = umma_smem_desc_sw128(smem_u32(a_blocks + (ks >> 2) * kBlkBytes + (ks & 3) * 32));
        const uint64_t db = umma_smem_desc_sw128(smem_u32(b_rows + ks * 2048));
        umma_f16(tmem_d, da, db, idesc, accumulate || ks != 0);
    }
}

// D[128, 64] (+)= A^T * B with A stored [16*ksteps rows (k) x 128 (m)] as two 64-column blocks (MN-major A) and
// B[16*ksteps x 64] (rows = k)
__device__ __forceinline__ void mma_rows_t_x_cols(uint32_t tmem_d, const uint8_t* a_rows, const uint8_t* b_rows,
                                                  int ksteps, bool accumulate) {
    const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = umma_smem_desc_sw128_lbo(smem_u32(a_rows + ks * 2048), kBlkBytes);
        const uint64_t db = umma_smem_desc_sw128(smem_u32(b_rows + ks * 2048));
        umma_f16(tmem_d, da, db, idesc, accumulate || ks != 0);
    }
}

// phase stamps of one CTA for tools/attn_trace.py; a null check per stamp when tracing is off
#define PCG_TRACE(slot)                                                                   \
    do {                                                                                  \
        if (p.trace != nullptr && lane == 0) p.trace[cta_id * 32 + (slot)] = clock64();   \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 192;  // warps 0-3 softmax (TMEM lane quarters), 4 TMA + MMA, 5 edge query row
constexpr int kFwdOffV = 2 * kBlkBytes;
constexpr int kFwdOffQ = 4 * kBlkBytes;
constexpr int kFwdOffX = 5 * kBlkBytes;                // float k_x[64], v_x[64], q_x[64], p_x[256]
constexpr int kFwdOffBar = kFwdOffX + 3 * 256 + 1024;  // 7 mbarriers + tmem slot
constexpr int kFwdSmemBytes = kFwdOffBar + 64 + 1024;
// TMEM columns: S [0, nk).  P (packed bf16 pairs) overwrites consumed S columns: keys >= 128 first, into
// [128, 192), then keys < 128 into [0, 64).  O accumulates in [192, 256), free once the first phase has read it.
constexpr uint32_t kFwdColPHi = 128, kFwdColO = 192;

struct FwdParams {
    int T, heads;
    int nv;  // T - 1: tokens on the tensor cores (as queries and as keys); token nv is the edge token
    int nk;  // nv rounded up to 16: MMA N of S / K extent of P V
    const bf16* qkv;
    bf16* out;
    float* lse;
    long long* trace;  // optional [ctas][32] clock64 stamps (tools/attn_trace.py), nullptr in production
    int stagger_ctas;    // the first wave: CTAs with a linear index below this ...
    int stagger_cycles;  // ... are delayed by this much when they are the second to arrive on their SM
};

// Two forward CTAs share an SM so that one's softmax (MUFU) overlaps the other's MMAs.  Launched together they
// run in lockstep and queue for the same unit at the same time; holding back every second arrival of the first
// wave by about half a CTA lifetime puts the pairs in anti-phase, and the offset then carries through the grid.
__device__ unsigned int g_fwd_sm_arrivals[1024];

__device__ __forceinline__ float chunk_max32(const uint32_t (&v)[32], int c, int nv, float mx) {
    float m0 = mx, m1 = -INFINITY;
    if (c + 32 <= nv) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            m0 = fmaxf(m0, __uint_as_float(v[j]));
            m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            m0 = fmaxf(m0, (c + j < nv) ? __uint_as_float(v[j]) : -INFINITY);
            m1 = fmaxf(m1, (c + j + 1 < nv) ? __uint_as_float(v[j + 1]) : -INFINITY);
        }
    }
    return fmaxf(m0, m1);
}
// row max over S columns [0, c_end), c_end a multiple of 32 (columns >= nv are masked, whatever they hold);
// the load of chunk c + 32 is in flight while chunk c is reduced
__device__ __forceinline__ float fwd_row_max(uint32_t trow, int c_end, int nv, float mx) {
    uint32_t va[32], vb[32];
    tmem_ld<32>(trow, va);
    for (int c = 0; c < c_end; c += 64) {
        tmem_wait_ld();
        if (c + 32 < c_end) tmem_ld<32>(trow + c + 32, vb);
        mx = chunk_max32(va, c, nv, mx);
        if (c + 32 < c_end) {
            tmem_wait_ld();
            if (c + 64 < c_end) tmem_ld<32>(trow + c + 64, va);
            mx = chunk_max32(vb, c + 32, nv, mx);
        }
    }
    return mx;
}
__device__ __forceinline__ float chunk_exp32(const uint32_t (&v)[32], int c, int nv, float mb, uint32_t tdst, float sum) {
    uint32_t pk[16];
    const bool full = c + 32 <= nv;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float e0 = exp2f(fmaf(__uint_as_float(v[2 * j]), kLog2e, -mb));
        float e1 = exp2f(fmaf(__uint_as_float(v[2 * j + 1]), kLog2e, -mb));
        if (!full) {
            e0 = (c + 2 * j < nv) ? e0 : 0.f;
            e1 = (c + 2 * j + 1 < nv) ? e1 : 0.f;
        }
        s0 += e0;
        s1 += e1;
        pk[j] = pack_bf16(e0, e1);
    }
    tmem_st<16>(tdst, pk);
    return sum + (s0 + s1);
}
// P = exp2(S log2e - mb) for S columns [c_begin, c_end) -> packed bf16 at TMEM columns p_col + (c - c_begin) / 2
__device__ __forceinline__ float fwd_row_exp(uint32_t trow, int c_begin, int c_end, uint32_t p_col, int nv, float mb,
                                             float sum) {
    uint32_t va[32], vb[32];
    tmem_ld<32>(trow + c_begin, va);
    for (int c = c_begin; c < c_end; c += 64) {
        tmem_wait_ld();
        if (c + 32 < c_end) tmem_ld<32>(trow + c + 32, vb);
        sum = chunk_exp32(va, c, nv, mb, trow + p_col + ((c - c_begin) >> 1), sum);
        if (c + 32 < c_end) {
            tmem_wait_ld();
            if (c + 64 < c_end) tmem_ld<32>(trow + c + 64, va);
            sum = chunk_exp32(vb, c + 32, nv, mb, trow + p_col + ((c + 32 - c_begin) >> 1), sum);
        }
    }
    return sum;
}

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const FwdParams p) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_k = sm;
    uint8_t* sm_v = sm + kFwdOffV;
    uint8_t* sm_q = sm + kFwdOffQ;
    float* kx = reinterpret_cast<float*>(sm + kFwdOffX);
    float* vx = kx + 64;
    float* qx = kx + 128;
    float* pbuf = kx + 192;  // [256] the edge row's probabilities
    // mbarriers: 0 K+Q landed, 1 V landed, 2 S ready, 3 P(keys >= 128) stored, 4 P(keys < 128) stored, 5 O ready,
    // 6 edge rows in shared memory
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kFwdOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int D = p.heads * kHd, nv = p.nv, nk = p.nk;
    const int q0 = tile * 128;
    const bool has_edge_row = (q0 + 128 >= nv);  // the CTA of the last tile also computes query row x
    const size_t cta_id = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (warp == 0) PCG_TRACE(0);

    uint32_t xq = 0, xk = 0, xv = 0;
    if (warp == 4) {
        if (lane == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            mbar_init(&bars[2], 1);
            mbar_init(&bars[3], 4);
            mbar_init(&bars[4], 4);
            mbar_init(&bars[5], 1);
            mbar_init(&bars[6], 1);
            fence_barrier_init();
            if (cta_id < static_cast<size_t>(p.stagger_ctas)) {
                uint32_t smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                if (atomicAdd(&g_fwd_sm_arrivals[smid & 1023], 1u) & 1u) {
                    const long long t0 = clock64();
                    while (clock64() - t0 < p.stagger_cycles) {
                    }
                }
            }
            // the loads only need the barriers: start them before the TMEM allocation and the CTA-wide sync
            const int nblk = (nk + 127) >> 7;
            mbar_arrive_expect_tx(&bars[0], (nblk + 1) * kBlkBytes);
            tma_load_3d(&map_qkv, &bars[0], sm_q, h * kHd, q0, n, kEvictFirst);
            for (int i = 0; i < nblk; ++i)
                tma_load_3d(&map_qkv, &bars[0], sm_k + i * kBlkBytes, D + h * kHd, i * 128, n, kEvictNormal);
            mbar_arrive_expect_tx(&bars[1], nblk * kBlkBytes);
            for (int i = 0; i < nblk; ++i)
                tma_load_3d(&map_qkv, &bars[1], sm_v + i * kBlkBytes, 2 * D + h * kHd, i * 128, n, kEvictNormal);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    } else if (warp == 5) {
        const bf16* xrow_g = p.qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + h * kHd + 2 * lane;
        xq = *reinterpret_cast<const uint32_t*>(xrow_g);
        xk = *reinterpret_cast<const uint32_t*>(xrow_g + D);
        xv = *reinterpret_cast<const uint32_t*>(xrow_g + 2 * D);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) PCG_TRACE(1);

    if (warp == 4) {
        if (lane == 0) {
            const int nblk = (nk + 127) >> 7;
            // pull the operands of the CTA that takes this slot next (stagger_ctas = CTAs resident at once) into L2;
            // the tiles of one head share K and V, so only the first tile's CTA prefetches those.  (Issued before this
            // CTA's own loads have landed: forward CTAs are short, later costs more lead time than it saves.)
            const size_t next = cta_id + p.stagger_ctas;
            if (next < static_cast<size_t>(gridDim.x) * gridDim.y * gridDim.z) {
                const int t2 = static_cast<int>(next % gridDim.x);
                const int h2 = static_cast<int>((next / gridDim.x) % gridDim.y), n2 = static_cast<int>(next / (gridDim.x * gridDim.y));
                tma_prefetch_3d(&map_qkv, h2 * kHd, t2 * 128, n2);
                if (t2 == 0)
                    for (int i = 0; i < nblk; ++i) {
                        tma_prefetch_3d(&map_qkv, D + h2 * kHd, i * 128, n2);
                        tma_prefetch_3d(&map_qkv, 2 * D + h2 * kHd, i * 128, n2);
                    }
            }
            mbar_wait(&bars[0], 0);
            PCG_TRACE(2);
            tc_fence_after();
            mma_tile_x_rows(tmem, sm_q, sm_k, nk);  // S = Q K^T
            umma_commit(&bars[2]);
            mbar_wait(&bars[1], 0);  // V landed
            // O = P V with A = P from TMEM (8 columns of bf16 pairs per 16 keys), keys >= 128 first
            const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
            const int ksteps = nk >> 4;
            if (ksteps > 8) {
                mbar_wait(&bars[3], 0);
                tc_fence_after();
                for (int ks = 8; ks < ksteps; ++ks)
                    umma_f16_ts(tmem + kFwdColO, tmem + kFwdColPHi + (ks - 8) * 8,
                                umma_smem_desc_sw128(smem_u32(sm_v + ks * 2048)), idesc, ks != 8);
            }
            mbar_wait(&bars[4], 0);
            tc_fence_after();
            for (int ks = 0; ks < min(ksteps, 8); ++ks)
                umma_f16_ts(tmem + kFwdColO, tmem + ks * 8, umma_smem_desc_sw128(smem_u32(sm_v + ks * 2048)), idesc,
                            ksteps > 8 || ks != 0);
            umma_commit(&bars[5]);
        }
    } else if (warp < 4) {
        const int r = warp * 32 + lane;  // query row in the tile == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        mbar_wait(&bars[6], 0);
        mbar_wait(&bars[0], 0);
        const float sx = row_dot(sm_q, r, kx);  // score against the edge key, while the MMA runs
        mbar_wait(&bars[2], 0);
        tc_fence_after();
        if (warp == 0) PCG_TRACE(4);
        const int nk32 = (nk + 31) & ~31;
        const float mx = fwd_row_max(trow, nk32, nv, sx);
        const float mb = mx * kLog2e;
        if (warp == 0) PCG_TRACE(5);
        float sum = 0.f;
        if (nk32 > 128) sum = fwd_row_exp(trow, 128, nk32, kFwdColPHi, nv, mb, sum);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[3]);
        sum = fwd_row_exp(trow, 0, min(nk32, 128), 0, nv, mb, sum);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[4]);
        if (warp == 0) PCG_TRACE(6);
        const float px = exp2f(fmaf(sx, kLog2e, -mb));
        sum += px;
        if (q0 + r < nv) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + q0 + r] = mx + logf(sum);
        const float inv = 1.0f / sum;
        mbar_wait(&bars[5], 0);
        tc_fence_after();
        if (warp == 0) PCG_TRACE(7);
        // O row: (P V + p_x v_x) / sum -> bf16 through the (now free) Q tile, then coalesced 16-byte stores
        uint32_t v[64];
        tmem_ld_32x32(trow + kFwdColO, reinterpret_cast<uint32_t(&)[32]>(v[0]));
        tmem_ld_32x32(trow + kFwdColO + 32, reinterpret_cast<uint32_t(&)[32]>(v[32]));
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float4 xa = *reinterpret_cast<const float4*>(vx + g * 8);
            const float4 xb = *reinterpret_cast<const float4*>(vx + g * 8 + 4);
            const uint4 o = make_uint4(pack_bf16(fmaf(px, xa.x, __uint_as_float(v[8 * g])) * inv,
                                                 fmaf(px, xa.y, __uint_as_float(v[8 * g + 1])) * inv),
                                       pack_bf16(fmaf(px, xa.z, __uint_as_float(v[8 * g + 2])) * inv,
                                                 fmaf(px, xa.w, __uint_as_float(v[8 * g + 3])) * inv),
                                       pack_bf16(fmaf(px, xb.x, __uint_as_float(v[8 * g + 4])) * inv,
                                                 fmaf(px, xb.y, __uint_as_float(v[8 * g + 5])) * inv),
                                       pack_bf16(fmaf(px, xb.z, __uint_as_float(v[8 * g + 6])) * inv,
                                                 fmaf(px, xb.w, __uint_as_float(v[8 * g + 7])) * inv));
            *reinterpret_cast<uint4*>(sm_q + row_chunk(r, g)) = o;
        }
        __syncwarp();
        bf16* gout = p.out + static_cast<size_t>(n) * p.T * D + h * kHd;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = warp * 32 + i * 4 + (lane >> 3), ch = lane & 7;
            if (q0 + row < nv)
                *reinterpret_cast<uint4*>(gout + static_cast<size_t>(q0 + row) * D + ch * 8) =
                    *reinterpret_cast<const uint4*>(sm_q + row_chunk(row, ch));
        }
        if (warp == 0) PCG_TRACE(8);
    } else {
        // edge token rows as float vectors for everyone (loaded into registers before the CTA-wide sync)
        qx[2 * lane] = bf_lo(xq), qx[2 * lane + 1] = bf_hi(xq);
        kx[2 * lane] = bf_lo(xk), kx[2 * lane + 1] = bf_hi(xk);
        vx[2 * lane] = bf_lo(xv), vx[2 * lane + 1] = bf_hi(xv);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[6]);
        if (has_edge_row) {
            // query row x against every key, on the CUDA cores: lane owns keys lane + 32 jj
            mbar_wait(&bars[0], 0);
            float s[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = lane + 32 * jj;
                s[jj] = (j < nv) ? row_dot(sm_k, j, qx) : -INFINITY;
            }
            float sxx = 0.f;
#pragma unroll
            for (int d = 0; d < 64; ++d) sxx = fmaf(qx[d], kx[d], sxx);
            float mx = sxx;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) mx = fmaxf(mx, s[jj]);
            mx = warp_max(mx);
            const float mb = mx * kLog2e;
            float part = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                s[jj] = exp2f(fmaf(s[jj], kLog2e, -mb));
                part += s[jj];
                pbuf[lane + 32 * jj] = s[jj];
            }
            const float exx = exp2f(fmaf(sxx, kLog2e, -mb));
            const float sum = warp_sum(part) + exx;
            __syncwarp();
            mbar_wait(&bars[1], 0);
            edge_gemv(pbuf, sm_v, nv, exx, vx, p.out + (static_cast<size_t>(n) * p.T + nv) * D + h * kHd, 1.0f / sum,
                      lane);
            if (lane == 0) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + nv] = mx + logf(sum);
            PCG_TRACE(9);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) PCG_TRACE(10);
    if (warp == 4) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// forward, persistent (the production kernel for 66 <= T <= 257)
// ---------------------------------------------------------------------------------------------------------
// Same arithmetic and the same tensor-memory plan as attn_fwd_tc_kernel, but a CTA no longer dies with its tile: the
// grid is 2 CTAs per SM and every CTA walks the work items (cutout, head, query tile) w = blockIdx.x + it * gridDim.x.
// What the one-shot kernel pays per tile -- barrier setup and the tensor-memory allocation (~1500 clk), the TMA round
// trip for Q, K and V (~1500 clk) and the drain of its last warp (~1000 clk) of a ~14000 clk lifetime -- is paid once
// per CTA, and the operands of item it + 1 are in flight while item it computes:
//   * Q is double buffered (slot it & 1; the slot also stages the O tile of its item for the coalesced stores),
//   * K is reloaded as soon as S(it) = Q K^T has retired and the edge warp has read it, V as soon as O(it) = P V has,
//   * S(it + 1) is issued the moment the softmax warps have pulled O(it) out of tensor memory, so it runs under
//     their global stores.
// One thread (warp 4, lane 0) issues both the TMA loads and the MMAs in a fixed order per item:
//   S(it) -> loads Q, K of it + 1 -> P V (it) -> load V of it + 1.
// Every mbarrier completes exactly once per item (or once per use of a Q slot) and no waiter can fall a phase behind:
// see the comments at the waits.
constexpr int kFwd2OffQ = 4 * kBlkBytes;                 // two Q slots
constexpr int kFwd2OffX = 6 * kBlkBytes;                 // 2 x {k_x[64], v_x[64], q_x[64]} floats, then p_x[256]
constexpr int kFwd2OffBar = kFwd2OffX + 2 * 768 + 1024;  // 16 mbarriers + tmem slot
constexpr int kFwd2SmemBytes = kFwd2OffBar + 16 * 8 + 64 + 1024;
enum {
    kB2FullQK = 0,    // [2] Q slot + K landed                      (TMA)
    kB2FullV = 2,     // V landed                                   (TMA)
    kB2SReady = 3,    // S = Q K^T retired                          (tcgen05.commit)
    kB2PHi = 4,       // P(keys >= 128) stored by the 4 softmax warps
    kB2PLo = 5,       // P(keys < 128) stored
    kB2OReady = 6,    // O = P V retired                            (tcgen05.commit)
    kB2TmemFree = 7,  // the 4 softmax warps have read O out of tensor memory
    kB2Done = 8,      // [2] the 4 softmax warps are done with Q slot / edge vectors of their item
    kB2XReady = 10,   // [2] edge-token vectors of the item written (edge warp)
    kB2EdgeK = 12,    // edge warp done reading K
    kB2EdgeV = 13,    // edge warp done reading V
    kB2Count = 14
};

// ---- compact softmax helpers of the persistent forward.  The one-shot kernel above unrolls a full and a masked
// variant of every 32-column step at every call site (~7k SASS instructions); run persistently, with two CTAs of six
// warps at different points of a 90 KB loop body, that code missed the instruction caches all the time (ncu: a quarter
// of all stall samples were no_instructions).  Here a partial chunk is patched to -inf in registers (32 selects, only
// for the last chunk of a short sequence) and a single variant of the arithmetic follows.
__device__ __noinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("pcg: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}
// bounded wait with the slow path out of line (the inline one costs ~30 instructions per call site)
__device__ __forceinline__ void mbar_wait_c(uint64_t* bar, uint32_t parity) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_spin(bar, parity);
}
__device__ __forceinline__ void mask_cols32(uint32_t (&v)[32], int c, int nv) {
    if (c + 32 > nv) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c + j >= nv) v[j] = 0xff800000u;  // -inf: drops out of the max, exp2 gives 0
    }
}
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float mx) {
    float m0 = mx, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        m0 = fmaxf(m0, __uint_as_float(v[j]));
        m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
    }
    return fmaxf(m0, m1);
}
__device__ __forceinline__ float exp32(const uint32_t (&v)[32], float mb, uint32_t tdst, float sum) {
    uint32_t pk[16];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float e0 = exp2f(fmaf(__uint_as_float(v[2 * j]), kLog2e, -mb));
        const float e1 = exp2f(fmaf(__uint_as_float(v[2 * j + 1]), kLog2e, -mb));
        s0 += e0;
        s1 += e1;
        pk[j] = pack_bf16(e0, e1);
    }
    tmem_st<16>(tdst, pk);
    return sum + (s0 + s1);
}

constexpr int kTraceItem = 5;  // the item of each CTA whose phases tools/attn_trace.py reports (steady state)

struct Fwd2Params {
    int T, heads;
    int nv, nk;
    int tiles;  // query tiles per (cutout, head)
    int items;  // n * heads * tiles
    const bf16* qkv;
    bf16* out;
    float* lse;
    long long* trace;    // optional [ctas][32] clock64 stamps of the CTA's second item
    int stagger_cycles;  // every second CTA to arrive on an SM starts this much later (anti-phase, see above)
};

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap map_qkv, const Fwd2Params p) {
    grid_dep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_k = sm;
    uint8_t* sm_v = sm + kFwdOffV;
    uint8_t* sm_q0 = sm + kFwd2OffQ;
    float* xvec = reinterpret_cast<float*>(sm + kFwd2OffX);  // slot s: k_x = xvec + 192 s, v_x = + 64, q_x = + 128
    float* pbuf = xvec + 384;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kFwd2OffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.heads * kHd, nv = p.nv, nk = p.nk;
    const int nblk = (nk + 127) >> 7;
    const int G = gridDim.x;
    const int n_my = (p.items - static_cast<int>(blockIdx.x) + G - 1) / G;
    const size_t cta_id = blockIdx.x;
    // item it of this CTA -> (cutout, head, first query row)
    auto decode = [&](int it, int& n, int& h, int& q0) {
        const int w = static_cast<int>(blockIdx.x) + it * G;
        const int nh = (p.tiles == 2) ? (w >> 1) : w;  // tiles is 1 or 2
        q0 = (p.tiles == 2) ? (w & 1) * 128 : 0;
        n = nh / p.heads;
        h = nh - n * p.heads;
    };
    auto load_qk = [&](int it) {
        int n, h, q0;
        decode(it, n, h, q0);
        uint64_t* bar = &bars[kB2FullQK + (it & 1)];
        mbar_arrive_expect_tx(bar, (nblk + 1) * kBlkBytes);
        tma_load_3d(&map_qkv, bar, sm_q0 + (it & 1) * kBlkBytes, h * kHd, q0, n, kEvictFirst);
        for (int i = 0; i < nblk; ++i)
            tma_load_3d(&map_qkv, bar, sm_k + i * kBlkBytes, D + h * kHd, i * 128, n, kEvictNormal);
    };
    auto load_v = [&](int it) {
        int n, h, q0;
        decode(it, n, h, q0);
        mbar_arrive_expect_tx(&bars[kB2FullV], nblk * kBlkBytes);
        for (int i = 0; i < nblk; ++i)
            tma_load_3d(&map_qkv, &bars[kB2FullV], sm_v + i * kBlkBytes, 2 * D + h * kHd, i * 128, n, kEvictNormal);
    };
    // the edge token's q, k, v rows of an item, two bf16 per lane (edge warp)
    auto load_x = [&](int it, uint32_t& xq, uint32_t& xk, uint32_t& xv) {
        int n, h, q0;
        decode(it, n, h, q0);
        const bf16* xrow_g = p.qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + h * kHd + 2 * lane;
        xq = *reinterpret_cast<const uint32_t*>(xrow_g);
        xk = *reinterpret_cast<const uint32_t*>(xrow_g + D);
        xv = *reinterpret_cast<const uint32_t*>(xrow_g + 2 * D);
    };

    uint32_t xq = 0, xk = 0, xv = 0;
    if (warp == 4) {
        if (lane == 0) {
            mbar_init(&bars[kB2FullQK], 1);
            mbar_init(&bars[kB2FullQK + 1], 1);
            mbar_init(&bars[kB2FullV], 1);
            mbar_init(&bars[kB2SReady], 1);
            mbar_init(&bars[kB2PHi], 4);
            mbar_init(&bars[kB2PLo], 4);
            mbar_init(&bars[kB2OReady], 1);
            mbar_init(&bars[kB2TmemFree], 4);
            mbar_init(&bars[kB2Done], 4);
            mbar_init(&bars[kB2Done + 1], 4);
            mbar_init(&bars[kB2XReady], 1);
            mbar_init(&bars[kB2XReady + 1], 1);
            mbar_init(&bars[kB2EdgeK], 1);
            mbar_init(&bars[kB2EdgeV], 1);
            fence_barrier_init();
            if (p.stagger_cycles > 0) {
                uint32_t smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                if (atomicAdd(&g_fwd_sm_arrivals[smid & 1023], 1u) & 1u) {
                    const long long t0 = clock64();
                    while (clock64() - t0 < p.stagger_cycles) {
                    }
                }
            }
            load_qk(0);
            load_v(0);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    } else if (warp == 5) {
        load_x(0, xq, xk, xv);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[cta_id * 32 + 30] = clock64(), p.trace[cta_id * 32 + 29] = n_my;

    if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
            const int ksteps = nk >> 4;
#pragma unroll 1
            for (int it = 0; it < n_my; ++it) {
                const int s = it & 1;
                const bool tr = (it == kTraceItem) && p.trace != nullptr;
                // ---- S(it) = Q K^T, as soon as the operands are in and O(it - 1) has left tensor memory
                mbar_wait_c(&bars[kB2FullQK + s], (it >> 1) & 1);
                if (tr) p.trace[cta_id * 32 + 10] = clock64();
                if (it > 0) mbar_wait_c(&bars[kB2TmemFree], (it - 1) & 1);
                if (tr) p.trace[cta_id * 32 + 11] = clock64();
                tc_fence_after();
                mma_tile_x_rows(tmem, sm_q0 + s * kBlkBytes, sm_k, nk);
                umma_commit(&bars[kB2SReady]);
                // ---- Q, K of item it + 1: the other Q slot is free once the epilogue of item it - 1 has drained it,
                // K once S(it) has retired and the edge warp has taken its dot products
                if (it + 1 < n_my) {
                    if (it >= 1) mbar_wait_c(&bars[kB2Done + (s ^ 1)], ((it - 1) >> 1) & 1);
                    mbar_wait_c(&bars[kB2SReady], it & 1);
                    mbar_wait_c(&bars[kB2EdgeK], it & 1);
                    load_qk(it + 1);
                }
                if (tr) p.trace[cta_id * 32 + 12] = clock64();
                // ---- O(it) = P V with A = P from tensor memory, keys >= 128 first
                mbar_wait_c(&bars[kB2FullV], it & 1);
                if (ksteps > 8) {
                    mbar_wait_c(&bars[kB2PHi], it & 1);
                    tc_fence_after();
                    for (int ks = 8; ks < ksteps; ++ks)
                        umma_f16_ts(tmem + kFwdColO, tmem + kFwdColPHi + (ks - 8) * 8,
                                    umma_smem_desc_sw128(smem_u32(sm_v + ks * 2048)), idesc_pv, ks != 8);
                }
                if (tr) p.trace[cta_id * 32 + 13] = clock64();
                mbar_wait_c(&bars[kB2PLo], it & 1);
                if (tr) p.trace[cta_id * 32 + 14] = clock64();
                tc_fence_after();
                for (int ks = 0; ks < min(ksteps, 8); ++ks)
                    umma_f16_ts(tmem + kFwdColO, tmem + ks * 8, umma_smem_desc_sw128(smem_u32(sm_v + ks * 2048)),
                                idesc_pv, ksteps > 8 || ks != 0);
                umma_commit(&bars[kB2OReady]);
                // ---- V of item it + 1 once P V has retired and the edge warp has finished its row
                if (it + 1 < n_my) {
                    mbar_wait_c(&bars[kB2OReady], it & 1);
                    if (tr) p.trace[cta_id * 32 + 15] = clock64();
                    mbar_wait_c(&bars[kB2EdgeV], it & 1);
                    load_v(it + 1);
                    if (tr) p.trace[cta_id * 32 + 16] = clock64();
                }
            }
        }
    } else if (warp < 4) {
        const int r = warp * 32 + lane;  // query row in the tile == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const int nch = (nk + 31) >> 5;       // 32-column chunks of S
        const int n_hi = max(nch - 4, 0);      // chunks of keys >= 128
        float sx_next = 0.f;
        bool have_sx = false;
#pragma unroll 1
        for (int it = 0; it < n_my; ++it) {
            const int s = it & 1;
            int n, h, q0;
            decode(it, n, h, q0);
            uint8_t* sm_q = sm_q0 + s * kBlkBytes;
            const float* kx = xvec + 192 * s;
            const float* vx = kx + 64;
            const bool tr = (it == kTraceItem) && warp == 0 && p.trace != nullptr && lane == 0;
            if (it == kTraceItem + 1 && warp == 0 && p.trace != nullptr && lane == 0) p.trace[cta_id * 32 + 9] = clock64();
            if (tr) p.trace[cta_id * 32 + 0] = clock64();
            // score against the edge key: normally taken one item ahead (below, while O = P V runs)
            float sx = sx_next;
            if (!have_sx) {
                mbar_wait_c(&bars[kB2XReady + s], (it >> 1) & 1);
                mbar_wait_c(&bars[kB2FullQK + s], (it >> 1) & 1);
                sx = row_dot(sm_q, r, kx);
            }
            have_sx = false;
            mbar_wait_c(&bars[kB2SReady], it & 1);
            tc_fence_after();
            if (tr) p.trace[cta_id * 32 + 4] = clock64();
            // row max over all S columns (32 at a time, the next load in flight behind the reduction)
            uint32_t va[32], vb[32];
            float mx = sx;
            tmem_ld<32>(trow, va);
#pragma unroll 1
            for (int i = 0; i < nch; i += 2) {
                tmem_wait_ld();
                if (i + 1 < nch) tmem_ld<32>(trow + 32 * (i + 1), vb);
                mask_cols32(va, 32 * i, nv);
                mx = max32(va, mx);
                if (i + 1 < nch) {
                    tmem_wait_ld();
                    if (i + 2 < nch) tmem_ld<32>(trow + 32 * (i + 2), va);
                    mask_cols32(vb, 32 * (i + 1), nv);
                    mx = max32(vb, mx);
                }
            }
            const float mb = mx * kLog2e;
            if (tr) p.trace[cta_id * 32 + 5] = clock64();
            // P = exp2(S log2e - mb) as packed bf16 over consumed S columns: the chunks of keys >= 128 first (into
            // columns [128, 192)), then keys < 128 (into [0, 64)); a store never reaches a column not yet loaded
            float sum = 0.f;
            auto col_of = [&](int i) { return i < n_hi ? 128 + 32 * i : 32 * (i - n_hi); };
            auto pcol_of = [&](int i) { return i < n_hi ? kFwdColPHi + 16 * i : 16 * (i - n_hi); };
            auto chunk_done = [&](int i) {
                if (i == n_hi - 1) {  // every key >= 128 is stored: the issuing thread may start O += P_hi V_hi
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[kB2PHi]);
                }
            };
            if (n_hi == 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[kB2PHi]);
            }
            tmem_ld<32>(trow + col_of(0), va);
#pragma unroll 1
            for (int i = 0; i < nch; i += 2) {
                tmem_wait_ld();
                if (i + 1 < nch) tmem_ld<32>(trow + col_of(i + 1), vb);
                mask_cols32(va, col_of(i), nv);
                sum = exp32(va, mb, trow + pcol_of(i), sum);
                chunk_done(i);
                if (i + 1 < nch) {
                    tmem_wait_ld();
                    if (i + 2 < nch) tmem_ld<32>(trow + col_of(i + 2), va);
                    mask_cols32(vb, col_of(i + 1), nv);
                    sum = exp32(vb, mb, trow + pcol_of(i + 1), sum);
                    chunk_done(i + 1);
                }
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2PLo]);
            if (tr) p.trace[cta_id * 32 + 6] = clock64();
            const float px = exp2f(fmaf(sx, kLog2e, -mb));
            sum += px;
            if (q0 + r < nv) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + q0 + r] = mx + logf(sum);
            const float inv = 1.0f / sum;
            // the tensor pipe needs ~1000 clk for O: take the next item's edge score now if its Q tile and edge vectors
            // have already landed (they usually have; both barriers stay complete until this warp's item it + 1)
            if (it + 1 < n_my && mbar_test_wait(&bars[kB2XReady + (s ^ 1)], ((it + 1) >> 1) & 1) &&
                mbar_test_wait(&bars[kB2FullQK + (s ^ 1)], ((it + 1) >> 1) & 1)) {
                sx_next = row_dot(sm_q0 + (s ^ 1) * kBlkBytes, r, xvec + 192 * (s ^ 1));
                have_sx = true;
            }
            mbar_wait_c(&bars[kB2OReady], it & 1);
            tc_fence_after();
            if (tr) p.trace[cta_id * 32 + 7] = clock64();
            uint32_t v[64];
            tmem_ld_32x32(trow + kFwdColO, reinterpret_cast<uint32_t(&)[32]>(v[0]));
            tmem_ld_32x32(trow + kFwdColO + 32, reinterpret_cast<uint32_t(&)[32]>(v[32]));
            tmem_wait_ld();
            // O is in registers: tensor memory may take S of the next item
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2TmemFree]);
            // Rows go straight to global memory from the tensor-memory layout (thread = row: eight 16-byte stores cover
            // the row's 128 contiguous bytes, one full line).  The one-shot kernel stages the tile through shared memory
            // for row-coalesced stores; here that round trip (store -> sync -> load -> store) sits on the per-item chain
            // and costs more than the extra store transactions.
            bf16* grow = p.out + (static_cast<size_t>(n) * p.T + q0 + r) * D + h * kHd;
            const bool row_ok = q0 + r < nv;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 xa = *reinterpret_cast<const float4*>(vx + g * 8);
                const float4 xb = *reinterpret_cast<const float4*>(vx + g * 8 + 4);
                const uint4 o = make_uint4(pack_bf16(fmaf(px, xa.x, __uint_as_float(v[8 * g])) * inv,
                                                     fmaf(px, xa.y, __uint_as_float(v[8 * g + 1])) * inv),
                                           pack_bf16(fmaf(px, xa.z, __uint_as_float(v[8 * g + 2])) * inv,
                                                     fmaf(px, xa.w, __uint_as_float(v[8 * g + 3])) * inv),
                                           pack_bf16(fmaf(px, xb.x, __uint_as_float(v[8 * g + 4])) * inv,
                                                     fmaf(px, xb.y, __uint_as_float(v[8 * g + 5])) * inv),
                                           pack_bf16(fmaf(px, xb.z, __uint_as_float(v[8 * g + 6])) * inv,
                                                     fmaf(px, xb.w, __uint_as_float(v[8 * g + 7])) * inv));
                if (row_ok) *reinterpret_cast<uint4*>(grow + g * 8) = o;
            }
            // this item's Q slot and edge vectors (v_x above was their last reader in this warp) may be refilled
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2Done + s]);
            if (tr) p.trace[cta_id * 32 + 8] = clock64();
        }
    } else {
        // edge warp: the edge token's vectors for everyone, and (for the CTA of the last tile) query row x on the CUDA
        // cores.  Its loads of item it + 1 are issued before the work of item it.
#pragma unroll 1
        for (int it = 0; it < n_my; ++it) {
            const int s = it & 1;
            int n, h, q0;
            decode(it, n, h, q0);
            const bool has_edge_row = (q0 + 128 >= nv);
            float* kx = xvec + 192 * s;
            float* vx = kx + 64;
            float* qx = kx + 128;
            // slot s was last read by the softmax warps of item it - 2
            if (it >= 2) mbar_wait_c(&bars[kB2Done + s], ((it - 2) >> 1) & 1);
            qx[2 * lane] = bf_lo(xq), qx[2 * lane + 1] = bf_hi(xq);
            kx[2 * lane] = bf_lo(xk), kx[2 * lane + 1] = bf_hi(xk);
            vx[2 * lane] = bf_lo(xv), vx[2 * lane + 1] = bf_hi(xv);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2XReady + s]);
            const float sxx = warp_sum(fmaf(bf_lo(xq), bf_lo(xk), bf_hi(xq) * bf_hi(xk)));  // q_x . k_x
            if (it + 1 < n_my) load_x(it + 1, xq, xk, xv);
            const bool tr = (it == kTraceItem) && p.trace != nullptr && lane == 0;
            if (tr) p.trace[cta_id * 32 + 20] = clock64();
            mbar_wait_c(&bars[kB2FullQK + s], (it >> 1) & 1);
            if (tr) p.trace[cta_id * 32 + 21] = clock64();
            float sc[8];
            if (has_edge_row) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = lane + 32 * jj;
                    sc[jj] = (j < nv) ? row_dot(sm_k, j, qx) : -INFINITY;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2EdgeK]);
            if (tr) p.trace[cta_id * 32 + 22] = clock64();
            if (has_edge_row) {
                float mx = sxx;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) mx = fmaxf(mx, sc[jj]);
                mx = warp_max(mx);
                const float mb = mx * kLog2e;
                float part = 0.f;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    sc[jj] = exp2f(fmaf(sc[jj], kLog2e, -mb));
                    part += sc[jj];
                    pbuf[lane + 32 * jj] = sc[jj];
                }
                const float exx = exp2f(fmaf(sxx, kLog2e, -mb));
                const float sum = warp_sum(part) + exx;
                __syncwarp();
                mbar_wait_c(&bars[kB2FullV], it & 1);
                edge_gemv(pbuf, sm_v, nv, exx, vx, p.out + (static_cast<size_t>(n) * p.T + nv) * D + h * kHd, 1.0f / sum,
                          lane);
                if (lane == 0) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + nv] = mx + logf(sum);
            } else {
                // no row to compute, but still keep step with the V loads: the arrival below must not run a phase
                // ahead of the issuing thread's wait for it (V of item it + 1 is only requested after that wait)
                mbar_wait_c(&bars[kB2FullV], it & 1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2EdgeV]);
            if (tr) p.trace[cta_id * 32 + 23] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[cta_id * 32 + 31] = clock64();
    if (warp == 4) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// forward, persistent, every score row split over two warps (production for 130 <= T <= 257: two key tiles)
// ---------------------------------------------------------------------------------------------------------
// The persistent kernel above walks one item in ~11 000 clk with FOUR softmax warps: thread = row, 256 scores per
// thread in sequence (row max 1400 clk, exponentials 4800, output 1900), so a (cutout, head, tile) item is a chain of
// dependent phases and the SM's MUFU and tensor pipes sit at a third of their rate even with two CTAs resident.  Here
// the same item is worked on by EIGHT softmax warps: warps 0-3 own the keys < 128 of their 32 rows (S columns [0, 128),
// P into [0, 64)), warps 4-7 the keys >= 128 (S columns [128, 256), P into [128, 192)); the two threads of a row
// exchange their partial row max (and the edge key's score) and later their partial sums through shared memory with a
// 64-thread named barrier per lane quarter, and each takes 32 of the 64 output columns.  Every phase of the chain is
// half as long, and four warps per scheduler instead of two keep the MUFU pipe fed.  Tensor memory, the operand
// ring, the MMA / TMA thread and the edge warp are those of attn_fwd_persist_kernel.
constexpr int kFwd3Threads = 320;  // warps 0-3 keys < 128, 4-7 keys >= 128, 8 TMA + MMA, 9 edge query row
constexpr int kFwd3OffEx = kFwd2OffX + 2 * 768 + 1024;  // float ex_mx[2][128], ex_sx[128], ex_sum[2][128]
constexpr int kFwd3OffBar = kFwd3OffEx + 5 * 512;
constexpr int kFwd3SmemBytes = kFwd3OffBar + 16 * 8 + 64 + 1024;

struct Fwd3Params : Fwd2Params {
    uint32_t heads_magic, grid_magic;  // ceil(2^32 / heads), ceil(2^32 / gridDim.x): exact quotients for the sizes the host admits
};

// Every kPolyEvery-th PAIR of exponentials is taken on the FMA / ALU pipes instead of MUFU (which is what the exp pass
// waits for: four softmax warps per scheduler, 16 ex2 per clock and SM): Cody-Waite split a = n + f with the round-to-
// nearest magic constant, 2^f on [-0.5, 0.5] as a cubic (relative error 7.5e-5; the result is rounded to bf16, 3.9e-3),
// 2^n added into the exponent field.  Packed fp32 arithmetic: 3 FMA-pipe + 2 ALU instructions per element.
#ifndef PCG_ATTN_POLY_EVERY
#define PCG_ATTN_POLY_EVERY 0
#endif
constexpr int kPolyEvery = PCG_ATTN_POLY_EVERY;
__device__ __forceinline__ float2 exp2_poly2(float2 a) {
    const float2 magic = make_float2(12582912.f, 12582912.f);
    a.x = fmaxf(a.x, -120.f), a.y = fmaxf(a.y, -120.f);  // -inf (masked column) and underflow: 2^-120 rounds to 0 in the sum
    const float2 t = __fadd2_rn(a, magic);
    const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = __fadd2_rn(a, make_float2(-n.x, -n.y));
    float2 p = __ffma2_rn(f, make_float2(0.05517165f, 0.05517165f), make_float2(0.24261112f, 0.24261112f));
    p = __ffma2_rn(p, f, make_float2(0.69326099f, 0.69326099f));
    p = __ffma2_rn(p, f, make_float2(0.99992807f, 0.99992807f));
    return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23)),
                       __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23)));
}

__device__ __forceinline__ float exp32_f2(const uint32_t (&v)[32], float mb, uint32_t tdst, float sum) {
    uint32_t pk[16];
    float2 s01 = make_float2(0.f, 0.f);
    const float2 l2 = make_float2(kLog2e, kLog2e), nmb = make_float2(-mb, -mb);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 a = __ffma2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), l2, nmb);
        const float2 e = (kPolyEvery > 0 && j % (kPolyEvery > 0 ? kPolyEvery : 1) == kPolyEvery - 1)
                             ? exp2_poly2(a)
                             : make_float2(exp2f(a.x), exp2f(a.y));
        s01 = __fadd2_rn(s01, e);
        pk[j] = pack_bf16(e.x, e.y);
    }
    tmem_st<16>(tdst, pk);
    return sum + (s01.x + s01.y);
}

__global__ void __launch_bounds__(kFwd3Threads, 2)
attn_fwd_split_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out,
                      const Fwd3Params p) {
    grid_dep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_k = sm;
    uint8_t* sm_v = sm + kFwdOffV;
    uint8_t* sm_q0 = sm + kFwd2OffQ;
    float* xvec = reinterpret_cast<float*>(sm + kFwd2OffX);  // slot s: k_x = xvec + 192 s, v_x = + 64, q_x = + 128
    float* pbuf = xvec + 384;
    float* ex_mx = reinterpret_cast<float*>(sm + kFwd3OffEx);  // [2][128]
    float* ex_sx = ex_mx + 256;                                // [128]
    float* ex_sum = ex_sx + 128;                               // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kFwd3OffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.heads * kHd, nv = p.nv, nk = p.nk;
    const int nblk = (nk + 127) >> 7;  // 2 here
    const int G = gridDim.x;
    // Item it of this CTA is w = it * G + (blockIdx.x + it) % G: round `it` of the grid is rotated by `it`, so that with
    // an even grid a CTA alternates between the two query tiles of a head.  Only the second tile carries the edge
    // query row, the one part of an item that a single warp works through on the CUDA cores (~9000 clk): with
    // w = blockIdx.x + it * G every odd CTA had it on every item and set the kernel's time.
    const int full_rounds = p.items / G;
    const int n_my = full_rounds + (((static_cast<int>(blockIdx.x) + full_rounds) % G) < p.items - full_rounds * G ? 1 : 0);
    // (no integer division here: the emulation's MUFU.RCP queues behind the exponentials of sixteen softmax warps, and
    // this sits between two items of every warp; the host passes ceil(2^32 / d) for both divisors)
    auto decode = [&](int it, int& n, int& h, int& q0) {
        const uint32_t x = blockIdx.x + static_cast<uint32_t>(it);
        const int w = it * G + static_cast<int>(x - __umulhi(x, p.grid_magic) * static_cast<uint32_t>(G));
        const int nh = w >> 1;  // two query tiles per (cutout, head)
        q0 = (w & 1) * 128;
        n = static_cast<int>(__umulhi(static_cast<uint32_t>(nh), p.heads_magic));
        h = nh - n * p.heads;
    };
    auto load_qk = [&](int it) {
        int n, h, q0;
        decode(it, n, h, q0);
        uint64_t* bar = &bars[kB2FullQK + (it & 1)];
        mbar_arrive_expect_tx(bar, (nblk + 1) * kBlkBytes);
        tma_load_3d(&map_qkv, bar, sm_q0 + (it & 1) * kBlkBytes, h * kHd, q0, n, kEvictFirst);
        for (int i = 0; i < nblk; ++i)
            tma_load_3d(&map_qkv, bar, sm_k + i * kBlkBytes, D + h * kHd, i * 128, n, kEvictNormal);
    };
    auto load_v = [&](int it) {
        int n, h, q0;
        decode(it, n, h, q0);
        mbar_arrive_expect_tx(&bars[kB2FullV], nblk * kBlkBytes);
        for (int i = 0; i < nblk; ++i)
            tma_load_3d(&map_qkv, &bars[kB2FullV], sm_v + i * kBlkBytes, 2 * D + h * kHd, i * 128, n, kEvictNormal);
    };
    auto load_x = [&](int it, uint32_t& xq, uint32_t& xk, uint32_t& xv) {
        int n, h, q0;
        decode(it, n, h, q0);
        const bf16* xrow_g = p.qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + h * kHd + 2 * lane;
        xq = *reinterpret_cast<const uint32_t*>(xrow_g);
        xk = *reinterpret_cast<const uint32_t*>(xrow_g + D);
        xv = *reinterpret_cast<const uint32_t*>(xrow_g + 2 * D);
    };

    uint32_t xq = 0, xk = 0, xv = 0;
    if (warp == 8) {
        if (lane == 0) {
            mbar_init(&bars[kB2FullQK], 1);
            mbar_init(&bars[kB2FullQK + 1], 1);
            mbar_init(&bars[kB2FullV], 1);
            mbar_init(&bars[kB2SReady], 1);
            mbar_init(&bars[kB2PHi], 4);
            mbar_init(&bars[kB2PLo], 4);
            mbar_init(&bars[kB2OReady], 1);
            mbar_init(&bars[kB2TmemFree], 8);
            mbar_init(&bars[kB2Done], 8);
            mbar_init(&bars[kB2Done + 1], 8);
            mbar_init(&bars[kB2XReady], 1);
            mbar_init(&bars[kB2XReady + 1], 1);
            mbar_init(&bars[kB2EdgeK], 1);
            mbar_init(&bars[kB2EdgeV], 1);
            fence_barrier_init();
            load_qk(0);
            load_v(0);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    } else if (warp == 9) {
        load_x(0, xq, xk, xv);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const size_t cta_id = blockIdx.x;
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[cta_id * 32 + 30] = clock64(), p.trace[cta_id * 32 + 29] = n_my;

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
            const int ksteps = nk >> 4;
#pragma unroll 1
            for (int it = 0; it < n_my; ++it) {
                const int s = it & 1;
                const bool tr = (it == kTraceItem) && p.trace != nullptr;
                // ---- S(it) = Q K^T, as soon as the operands are in and O(it - 1) has left tensor memory
                mbar_wait_c(&bars[kB2FullQK + s], (it >> 1) & 1);
                if (tr) p.trace[cta_id * 32 + 10] = clock64();
                if (it > 0) mbar_wait_c(&bars[kB2TmemFree], (it - 1) & 1);
                if (tr) p.trace[cta_id * 32 + 11] = clock64();
                tc_fence_after();
                mma_tile_x_rows(tmem, sm_q0 + s * kBlkBytes, sm_k, nk);
                umma_commit(&bars[kB2SReady]);
                if (it + 1 < n_my) {
                    if (it >= 1) mbar_wait_c(&bars[kB2Done + (s ^ 1)], ((it - 1) >> 1) & 1);
                    mbar_wait_c(&bars[kB2SReady], it & 1);
                    mbar_wait_c(&bars[kB2EdgeK], it & 1);
                    load_qk(it + 1);
                }
                if (tr) p.trace[cta_id * 32 + 12] = clock64();
                // ---- O(it) = P V with A = P from tensor memory, keys >= 128 first
                mbar_wait_c(&bars[kB2FullV], it & 1);
                mbar_wait_c(&bars[kB2PHi], it & 1);
                if (tr) p.trace[cta_id * 32 + 13] = clock64();
                tc_fence_after();
                for (int ks = 8; ks < ksteps; ++ks)
                    umma_f16_ts(tmem + kFwdColO, tmem + kFwdColPHi + (ks - 8) * 8,
                                umma_smem_desc_sw128(smem_u32(sm_v + ks * 2048)), idesc_pv, ks != 8);
                mbar_wait_c(&bars[kB2PLo], it & 1);
                if (tr) p.trace[cta_id * 32 + 14] = clock64();
                tc_fence_after();
                for (int ks = 0; ks < 8; ++ks)
                    umma_f16_ts(tmem + kFwdColO, tmem + ks * 8, umma_smem_desc_sw128(smem_u32(sm_v + ks * 2048)),
                                idesc_pv, 1);
                umma_commit(&bars[kB2OReady]);
                if (it + 1 < n_my) {
                    mbar_wait_c(&bars[kB2OReady], it & 1);
                    if (tr) p.trace[cta_id * 32 + 15] = clock64();
                    mbar_wait_c(&bars[kB2EdgeV], it & 1);
                    load_v(it + 1);
                    if (tr) p.trace[cta_id * 32 + 16] = clock64();
                }
            }
        }
    } else if (warp < 8) {
        const int grp = warp >> 2, quarter = warp & 3;  // grp 0: keys < 128, grp 1: keys >= 128
        const int r = quarter * 32 + lane;              // query row in the tile == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        const int nch = (nk + 31) >> 5;                       // 32-column chunks of S (5..8)
        const int my_ch = grp == 0 ? 4 : nch - 4;             // chunks of this group
        const uint32_t s_col = grp == 0 ? 0u : 128u;          // first S column of this group
        const uint32_t p_col = grp == 0 ? 0u : kFwdColPHi;    // first packed P column of this group
        float sx_next = 0.f;
        bool have_sx = false;
#pragma unroll 1
        for (int it = 0; it < n_my; ++it) {
            const int s = it & 1;
            int n, h, q0;
            decode(it, n, h, q0);
            const float* kx = xvec + 192 * s;
            const float* vx = kx + 64;
            const bool tr = (it == kTraceItem) && quarter == 0 && p.trace != nullptr && lane == 0;
            const int tb = grp == 0 ? 0 : 17 - 4;  // group 1 stamps S ready / exp done / epilogue done at 17 / 18 / 19
            if (it == kTraceItem + 1 && warp == 0 && p.trace != nullptr && lane == 0) p.trace[cta_id * 32 + 9] = clock64();
            if (tr && grp == 0) p.trace[cta_id * 32 + 0] = clock64(), p.trace[cta_id * 32 + 28] = 1 + (q0 + 128 >= nv);
            // score against the edge key (group 0; normally taken one item ahead, below)
            float sx = sx_next;
            if (grp == 0 && !have_sx) {
                mbar_wait_c(&bars[kB2XReady + s], (it >> 1) & 1);
                mbar_wait_c(&bars[kB2FullQK + s], (it >> 1) & 1);
                sx = row_dot(sm_q0 + s * kBlkBytes, r, kx);
            }
            have_sx = false;
            mbar_wait_c(&bars[kB2SReady], it & 1);
            tc_fence_after();
            if (tr) p.trace[cta_id * 32 + tb + 4] = clock64();
            uint32_t va[32];
            float mx = grp == 0 ? sx : -INFINITY;
#pragma unroll 1
            for (int i = 0; i < my_ch; ++i) {
                tmem_ld<32>(trow + s_col + 32 * i, va);
                tmem_wait_ld();
                mask_cols32(va, static_cast<int>(s_col) + 32 * i, nv);
                mx = max32(va, mx);
            }
            // the row's other half
            ex_mx[grp * 128 + r] = mx;
            if (grp == 0) ex_sx[r] = sx;
            named_bar_sync(1 + quarter, 64);
            mx = fmaxf(mx, ex_mx[(grp ^ 1) * 128 + r]);
            sx = ex_sx[r];
            if (tr && grp == 0) p.trace[cta_id * 32 + 5] = clock64();
            const float mb = mx * kLog2e;
            float sum = 0.f;
#pragma unroll 1
            for (int i = 0; i < my_ch; ++i) {
                tmem_ld<32>(trow + s_col + 32 * i, va);
                tmem_wait_ld();
                mask_cols32(va, static_cast<int>(s_col) + 32 * i, nv);
                sum = exp32_f2(va, mb, trow + p_col + 16 * i, sum);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[grp == 0 ? kB2PLo : kB2PHi]);
            if (tr) p.trace[cta_id * 32 + (grp == 0 ? 6 : 18)] = clock64();
            const float px = exp2f(fmaf(sx, kLog2e, -mb));
            ex_sum[grp * 128 + r] = sum;
            named_bar_sync(1 + quarter, 64);
            sum += ex_sum[(grp ^ 1) * 128 + r] + px;
            if (grp == 0 && q0 + r < nv) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + q0 + r] = mx + logf(sum);
            const float inv = 1.0f / sum;
            // the tensor pipe needs ~1500 clk for O: group 0 takes the next item's edge score now if its Q tile and
            // edge vectors have landed (both barriers stay complete until this warp's item it + 1)
            if (grp == 0 && it + 1 < n_my && mbar_test_wait(&bars[kB2XReady + (s ^ 1)], ((it + 1) >> 1) & 1) &&
                mbar_test_wait(&bars[kB2FullQK + (s ^ 1)], ((it + 1) >> 1) & 1)) {
                sx_next = row_dot(sm_q0 + (s ^ 1) * kBlkBytes, r, xvec + 192 * (s ^ 1));
                have_sx = true;
            }
            mbar_wait_c(&bars[kB2OReady], it & 1);
            tc_fence_after();
            if (tr && grp == 0) p.trace[cta_id * 32 + 7] = clock64();
            tmem_ld<32>(trow + kFwdColO + 32 * grp, va);
            tmem_wait_ld();
            // this thread's half of O is in registers: tensor memory may take S of the next item
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2TmemFree]);
            // The O tile goes out through this item's Q slot (dead: S has retired and every warp has taken its edge score)
            // as a 128-byte-swizzled [128 x 64] bf16 tile and one bulk tensor store per lane quarter.  Rows stored
            // straight from the tensor-memory layout cost 32 L1 wavefronts per store instruction, which the other
            // resident CTA's warps queue behind; rows >= nv are clipped by the store map.
            uint8_t* stage = sm_q0 + s * kBlkBytes;
            const float* vxg = vx + 32 * grp;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 xa = *reinterpret_cast<const float4*>(vxg + g * 8);
                const float4 xb = *reinterpret_cast<const float4*>(vxg + g * 8 + 4);
                const uint4 o = make_uint4(pack_bf16(fmaf(px, xa.x, __uint_as_float(va[8 * g])) * inv,
                                                     fmaf(px, xa.y, __uint_as_float(va[8 * g + 1])) * inv),
                                           pack_bf16(fmaf(px, xa.z, __uint_as_float(va[8 * g + 2])) * inv,
                                                     fmaf(px, xa.w, __uint_as_float(va[8 * g + 3])) * inv),
                                           pack_bf16(fmaf(px, xb.x, __uint_as_float(va[8 * g + 4])) * inv,
                                                     fmaf(px, xb.y, __uint_as_float(va[8 * g + 5])) * inv),
                                           pack_bf16(fmaf(px, xb.z, __uint_as_float(va[8 * g + 6])) * inv,
                                                     fmaf(px, xb.w, __uint_as_float(va[8 * g + 7])) * inv));
                *reinterpret_cast<uint4*>(stage + row_chunk(r, 4 * grp + g)) = o;
            }
            fence_proxy_async();
            named_bar_sync(1 + quarter, 64);  // both column halves of rows [32 quarter, +32) are in the tile
            if (grp == 0) {
                if (lane == 0) {
                    tma_store_3d(&map_out, stage + quarter * 4096, h * kHd, q0 + 32 * quarter, n);
                    bulk_commit_group();
                    bulk_wait_read_all();  // the slot may be refilled once the store has read it
                }
                __syncwarp();
            }
            // this item's Q slot and edge vectors (v_x above was their last reader in this warp) may be refilled
            if (lane == 0) mbar_arrive(&bars[kB2Done + s]);
            if (tr) p.trace[cta_id * 32 + (grp == 0 ? 8 : 19)] = clock64();
        }
        if (grp == 0 && lane == 0) bulk_wait_all();
    } else {
        // edge warp: as in attn_fwd_persist_kernel
#pragma unroll 1
        for (int it = 0; it < n_my; ++it) {
            const int s = it & 1;
            int n, h, q0;
            decode(it, n, h, q0);
            const bool has_edge_row = (q0 + 128 >= nv);
            float* kx = xvec + 192 * s;
            float* vx = kx + 64;
            float* qx = kx + 128;
            if (it >= 2) mbar_wait_c(&bars[kB2Done + s], ((it - 2) >> 1) & 1);
            qx[2 * lane] = bf_lo(xq), qx[2 * lane + 1] = bf_hi(xq);
            kx[2 * lane] = bf_lo(xk), kx[2 * lane + 1] = bf_hi(xk);
            vx[2 * lane] = bf_lo(xv), vx[2 * lane + 1] = bf_hi(xv);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2XReady + s]);
            const float sxx = warp_sum(fmaf(bf_lo(xq), bf_lo(xk), bf_hi(xq) * bf_hi(xk)));  // q_x . k_x
            if (it + 1 < n_my) load_x(it + 1, xq, xk, xv);
            const bool tr = (it == kTraceItem) && p.trace != nullptr && lane == 0;
            if (tr) p.trace[cta_id * 32 + 20] = clock64();
            mbar_wait_c(&bars[kB2FullQK + s], (it >> 1) & 1);
            if (tr) p.trace[cta_id * 32 + 21] = clock64();
            float sc[8];
            if (has_edge_row) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = lane + 32 * jj;
                    sc[jj] = (j < nv) ? row_dot(sm_k, j, qx) : -INFINITY;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2EdgeK]);
            if (tr) p.trace[cta_id * 32 + 22] = clock64();
            if (has_edge_row) {
                float mx = sxx;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) mx = fmaxf(mx, sc[jj]);
                mx = warp_max(mx);
                const float mb = mx * kLog2e;
                float part = 0.f;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    sc[jj] = exp2f(fmaf(sc[jj], kLog2e, -mb));
                    part += sc[jj];
                    pbuf[lane + 32 * jj] = sc[jj];
                }
                const float exx = exp2f(fmaf(sxx, kLog2e, -mb));
                const float sum = warp_sum(part) + exx;
                __syncwarp();
                mbar_wait_c(&bars[kB2FullV], it & 1);
                edge_gemv(pbuf, sm_v, nv, exx, vx, p.out + (static_cast<size_t>(n) * p.T + nv) * D + h * kHd, 1.0f / sum,
                          lane);
                if (lane == 0) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + nv] = mx + logf(sum);
            } else {
                mbar_wait_c(&bars[kB2FullV], it & 1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kB2EdgeV]);
            if (tr) p.trace[cta_id * 32 + 23] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[cta_id * 32 + 31] = clock64();
    if (warp == 8) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// forward, long sequences (T - 1 > 256: ViT-L/14 @336 has 576 + 1 tokens)
// ---------------------------------------------------------------------------------------------------------
// Same CTA shape as above (cutout, head, 128-query tile; 2 CTAs / SM), but S no longer fits tensor memory, so the
// keys stream through a two-stage ring of 128-row K / V blocks with the online softmax: per chunk
//   S_c = Q K_c^T  ->  m' = max(m, rowmax S_c),  P_c = exp2(S_c - m') over the dead S columns,
//   l = l alpha + rowsum P_c,  O = O alpha (tcgen05.ld / st on the 64 accumulator columns, alpha = exp2(m - m'))
//   ->  O += P_c V_c (A = P_c from TMEM).
// The MMAs of one chunk and the softmax of the same chunk are serial inside a CTA; the second CTA of the SM fills
// the gaps.  There is no edge token here: with five or more tiles the class token simply rides in the last,
// partly filled tile (577 = 4 x 128 + 65), p.nv = T.
constexpr int kFlashOffBar = kFwdOffX;  // 8 mbarriers + tmem slot
constexpr int kFlashSmemBytes = kFlashOffBar + 128 + 1024;
constexpr uint32_t kFlashColO = 128;
enum { kFbK = 0, kFbV = 2, kFbS = 4, kFbP = 5, kFbO = 6, kFbX = 7 };

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_flash_kernel(const __grid_constant__ CUtensorMap map_qkv, const FwdParams p) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_k = sm;             // two stages of [128 keys x 64]
    uint8_t* sm_v = sm + kFwdOffV;  // two stages
    uint8_t* sm_q = sm + kFwdOffQ;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kFlashOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int D = p.heads * kHd, nv = p.nv;
    const int q0 = tile * 128;
    const int nc = (nv + 127) >> 7;
    const size_t cta_id = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;

    if (warp == 4) {
        if (lane == 0) {
            mbar_init(&bars[kFbK], 1);
            mbar_init(&bars[kFbK + 1], 1);
            mbar_init(&bars[kFbV], 1);
            mbar_init(&bars[kFbV + 1], 1);
            mbar_init(&bars[kFbS], 1);
            mbar_init(&bars[kFbP], 4);
            mbar_init(&bars[kFbO], 1);
            mbar_init(&bars[kFbX], 1);
            fence_barrier_init();
            if (cta_id < static_cast<size_t>(p.stagger_ctas)) {
                uint32_t smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                if (atomicAdd(&g_fwd_sm_arrivals[smid & 1023], 1u) & 1u) {
                    const long long t0 = clock64();
                    while (clock64() - t0 < p.stagger_cycles) {
                    }
                }
            }
            mbar_arrive_expect_tx(&bars[kFbK], 2 * kBlkBytes);
            tma_load_3d(&map_qkv, &bars[kFbK], sm_q, h * kHd, q0, n, kEvictFirst);
            tma_load_3d(&map_qkv, &bars[kFbK], sm_k, D + h * kHd, 0, n, kEvictNormal);
            mbar_arrive_expect_tx(&bars[kFbV], kBlkBytes);
            tma_load_3d(&map_qkv, &bars[kFbV], sm_v, 2 * D + h * kHd, 0, n, kEvictNormal);
            if (nc > 1) {
                mbar_arrive_expect_tx(&bars[kFbK + 1], kBlkBytes);
                tma_load_3d(&map_qkv, &bars[kFbK + 1], sm_k + kBlkBytes, D + h * kHd, 128, n, kEvictNormal);
                mbar_arrive_expect_tx(&bars[kFbV + 1], kBlkBytes);
                tma_load_3d(&map_qkv, &bars[kFbV + 1], sm_v + kBlkBytes, 2 * D + h * kHd, 128, n, kEvictNormal);
            }
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
            for (int c = 0; c < nc; ++c) {
                const int st = c & 1;
                const uint32_t ph = (c >> 1) & 1;
                mbar_wait(&bars[kFbK + st], ph);
                tc_fence_after();
                // S_c (it queues behind P_{c-1} V_{c-1}, which reads the columns it overwrites: same pipe, in order)
                mma_tile_x_rows(tmem, sm_q, sm_k + st * kBlkBytes, 128);
                umma_commit(&bars[kFbS]);
                if (c >= 1 && c + 1 < nc) {
                    // the other stage held chunk c - 1: its K block was consumed by S_{c-1} (done before P_{c-1} was
                    // signalled), its V block is free once P_{c-1} V_{c-1} has retired
                    const int s2 = st ^ 1;
                    mbar_arrive_expect_tx(&bars[kFbK + s2], kBlkBytes);
                    tma_load_3d(&map_qkv, &bars[kFbK + s2], sm_k + s2 * kBlkBytes, D + h * kHd, (c + 1) * 128, n, kEvictNormal);
                    mbar_wait(&bars[kFbO], (c - 1) & 1);
                    mbar_arrive_expect_tx(&bars[kFbV + s2], kBlkBytes);
                    tma_load_3d(&map_qkv, &bars[kFbV + s2], sm_v + s2 * kBlkBytes, 2 * D + h * kHd, (c + 1) * 128, n, kEvictNormal);
                }
                mbar_wait(&bars[kFbV + st], ph);
                mbar_wait(&bars[kFbP], c & 1);
                tc_fence_after();
                const int ksteps = min(8, (nv - c * 128 + 15) >> 4);
                for (int ks = 0; ks < ksteps; ++ks)
                    umma_f16_ts(tmem + kFlashColO, tmem + ks * 8,
                                umma_smem_desc_sw128(smem_u32(sm_v + st * kBlkBytes + ks * 2048)), idesc_pv, c > 0 || ks != 0);
                umma_commit(&bars[kFbO]);
            }
        }
    } else if (warp < 4) {
        const int r = warp * 32 + lane;
        const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        float mx = -INFINITY, sum = 0.f;
        for (int c = 0; c < nc; ++c) {
            const int nvl = min(128, nv - c * 128);  // valid keys of this chunk
            const int cend = (nvl + 31) & ~31;
            mbar_wait(&bars[kFbS], c & 1);
            tc_fence_after();
            const float m_new = fwd_row_max(trow, cend, nvl, mx);
            const float alpha = exp2f((mx - m_new) * kLog2e);
            mx = m_new;
            sum = fwd_row_exp(trow, 0, cend, 0, nvl, m_new * kLog2e, sum * alpha);
            if (c > 0) {
                mbar_wait(&bars[kFbO], (c - 1) & 1);  // O holds chunks 0 .. c-1 (already true: S_c was issued after them)
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t o[32];
                        tmem_ld<32>(trow + kFlashColO + 32 * hh, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                        tmem_st<16>(trow + kFlashColO + 32 * hh, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                        tmem_st<16>(trow + kFlashColO + 32 * hh + 16, reinterpret_cast<uint32_t(&)[16]>(o[16]));
                    }
                }
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kFbP]);
        }
        if (q0 + r < nv) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + q0 + r] = mx + logf(sum);
        const float inv = 1.0f / sum;
        mbar_wait(&bars[kFbO], (nc - 1) & 1);
        tc_fence_after();
        uint32_t v[64];
        tmem_ld_32x32(trow + kFlashColO, reinterpret_cast<uint32_t(&)[32]>(v[0]));
        tmem_ld_32x32(trow + kFlashColO + 32, reinterpret_cast<uint32_t(&)[32]>(v[32]));
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const uint4 o = make_uint4(pack_bf16(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv),
                                       pack_bf16(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv),
                                       pack_bf16(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv),
                                       pack_bf16(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv));
            *reinterpret_cast<uint4*>(sm_q + row_chunk(r, g)) = o;
        }
        __syncwarp();
        bf16* gout = p.out + static_cast<size_t>(n) * p.T * D + h * kHd;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = warp * 32 + i * 4 + (lane >> 3), ch = lane & 7;
            if (q0 + row < nv)
                *reinterpret_cast<uint4*>(gout + static_cast<size_t>(q0 + row) * D + ch * 8) =
                    *reinterpret_cast<const uint4*>(sm_q + row_chunk(row, ch));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 384;  // warps 0-7 elementwise (lane quarter = warp & 3, column phase = warp >> 2),
                                  // warp 8 TMA + MMA, warps 9-11 edge token (row x and column x of the scores)
constexpr int kBwdOffQ = 0;
constexpr int kBwdOffK = 2 * kBlkBytes;
constexpr int kBwdOffV = 4 * kBlkBytes;
constexpr int kBwdOffDO = 6 * kBlkBytes;
// dS^T [128 keys x 128 queries] bf16 as two 64-column blocks, TWO buffers (block parity): dQ(b) reads one while the
// elementwise warps fill the other for block b + 1.  P^T does not live in shared memory: it is written back over the
// S^T columns of tensor memory as packed bf16 and read by dV = P^T dO as a TMEM A operand.
constexpr int kBwdOffDST = 8 * kBlkBytes;
constexpr int kBwdOffStage = 12 * kBlkBytes;  // 8 warps x [32 rows x 64 B] epilogue staging
constexpr int kBwdOffVec = 13 * kBlkBytes;    // float[256] x 6: lse2, delta, pcol, dscol, prow, dsrow
constexpr int kBwdOffX = kBwdOffVec + 6 * 1024;  // float[64] x 4: q_x, k_x, v_x, dO_x; then 8 scalars
constexpr int kBwdOffBar = kBwdOffX + 4 * 256 + 32;
constexpr int kBwdSmemBytes = kBwdOffBar + 128 + 1024;
constexpr uint32_t kColST = 0, kColDPT = 128, kColDV = 256, kColDK = 320, kColDQ = 384;

struct BwdParams {
    int T, heads;
    int nv;       // T - 1
    int n_tiles;  // ceil(nv / 128): 1 or 2
    const bf16* qkv;
    const bf16* out;
    const bf16* d_out;
    const float* lse;
    bf16* d_qkv;
    long long* trace;
    int wave;  // CTAs resident at once (one per SM): the L2 prefetch distance
    int stagger_cycles;  // first wave: SM s starts (s % 8) * stagger_cycles late, see the kernel
};

// 16 columns of one block for this thread's key row: P^T = exp2(S^T log2e - lse2), dS^T = P^T (dP^T - delta) as
// packed bf16 (two 16-byte chunks each) ...
struct Cols16 {
    uint4 p[2], ds[2];
};
__device__ __forceinline__ Cols16 bwd_cols16(const uint32_t (&s)[16], const uint32_t (&dp)[16], const float* lse2,
                                             const float* delta, bool row_ok) {
    Cols16 o;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const float4 la = *reinterpret_cast<const float4*>(lse2 + 8 * g);
        const float4 lb = *reinterpret_cast<const float4*>(lse2 + 8 * g + 4);
        const float4 da = *reinterpret_cast<const float4*>(delta + 8 * g);
        const float4 db = *reinterpret_cast<const float4*>(delta + 8 * g + 4);
        const float l[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
        const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        float pv[8], dsv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            pv[k] = exp2f(fmaf(__uint_as_float(s[8 * g + k]), kLog2e, -l[k]));
            dsv[k] = pv[k] * (__uint_as_float(dp[8 * g + k]) - d[k]);
        }
        o.p[g] = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                            pack_bf16(pv[6], pv[7]));
        o.ds[g] = make_uint4(pack_bf16(dsv[0], dsv[1]), pack_bf16(dsv[2], dsv[3]), pack_bf16(dsv[4], dsv[5]),
                             pack_bf16(dsv[6], dsv[7]));
        if (!row_ok) o.p[g] = o.ds[g] = make_uint4(0u, 0u, 0u, 0u);
    }
    return o;
}
// ... and their stores into the swizzled [128 x 64] shared-memory blocks (16-byte chunks chunk0, chunk0 + 1 of row r)
__device__ __forceinline__ void bwd_store16(const Cols16& o, uint8_t* pt_blk, uint8_t* dst_blk, int r, int chunk0) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const uint32_t off = row_chunk(r, chunk0 + g);
        *reinterpret_cast<uint4*>(pt_blk + off) = o.p[g];
        *reinterpret_cast<uint4*>(dst_blk + off) = o.ds[g];
    }
}

// dS^T only (the packed kernels keep P^T in tensor memory)
__device__ __forceinline__ void bwd_store_ds(const Cols16& o, uint8_t* dst_blk, int r, int chunk0) {
#pragma unroll
    for (int g = 0; g < 2; ++g) *reinterpret_cast<uint4*>(dst_blk + row_chunk(r, chunk0 + g)) = o.ds[g];
}

// 32 accumulator columns of this thread's row (+ coef * xrow[.] when kEdge) -> bf16 -> the warp's staging rows -> global
template <bool kEdge = true>
__device__ __forceinline__ void bwd_epilogue(uint32_t taddr, float coef, const float* xrow, uint8_t* stage, int lane,
                                             bf16* gbase, size_t ld, int row0, int row_end) {
    uint32_t v[32];
    tmem_ld<32>(taddr, v);
    tmem_wait_ld();
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;
        if constexpr (kEdge) {
            xa = *reinterpret_cast<const float4*>(xrow + g * 8);
            xb = *reinterpret_cast<const float4*>(xrow + g * 8 + 4);
        }
        const uint4 o = make_uint4(pack_bf16(fmaf(coef, xa.x, __uint_as_float(v[8 * g])),
                                             fmaf(coef, xa.y, __uint_as_float(v[8 * g + 1]))),
                                   pack_bf16(fmaf(coef, xa.z, __uint_as_float(v[8 * g + 2])),
                                             fmaf(coef, xa.w, __uint_as_float(v[8 * g + 3]))),
                                   pack_bf16(fmaf(coef, xb.x, __uint_as_float(v[8 * g + 4])),
                                             fmaf(coef, xb.y, __uint_as_float(v[8 * g + 5]))),
                                   pack_bf16(fmaf(coef, xb.z, __uint_as_float(v[8 * g + 6])),
                                             fmaf(coef, xb.w, __uint_as_float(v[8 * g + 7]))));
        *reinterpret_cast<uint4*>(stage + lane * 64 + ((g ^ sw) << 4)) = o;
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = it * 32 + lane, row = idx >> 2, ch = idx & 3;
        if (row0 + row < row_end)
            *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(row0 + row) * ld + ch * 8) =
                *reinterpret_cast<const uint4*>(stage + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
    }
    __syncwarp();
}

// partial sum_d a[d] * b[d] over one 16-byte chunk (8 bf16) of two rows
__device__ __forceinline__ float chunk_dot(const uint4& x, const uint4& y) {
    float acc0 = bf_lo(x.x) * bf_lo(y.x), acc1 = bf_hi(x.x) * bf_hi(y.x);
    acc0 = fmaf(bf_lo(x.y), bf_lo(y.y), acc0);
    acc1 = fmaf(bf_hi(x.y), bf_hi(y.y), acc1);
    acc0 = fmaf(bf_lo(x.z), bf_lo(y.z), acc0);
    acc1 = fmaf(bf_hi(x.z), bf_hi(y.z), acc1);
    acc0 = fmaf(bf_lo(x.w), bf_lo(y.w), acc0);
    acc1 = fmaf(bf_hi(x.w), bf_hi(y.w), acc1);
    return acc0 + acc1;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                   const BwdParams p) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_q = sm + kBwdOffQ;
    uint8_t* sm_k = sm + kBwdOffK;
    uint8_t* sm_v = sm + kBwdOffV;
    uint8_t* sm_do = sm + kBwdOffDO;
    uint8_t* sm_dst = sm + kBwdOffDST;
    float* lse2 = reinterpret_cast<float*>(sm + kBwdOffVec);  // lse * log2(e), +inf past the last tensor-core query
    float* delta = lse2 + 256;   // rowsum(dO * O)
    float* pcol = lse2 + 512;    // p(query i, key x)
    float* dscol = lse2 + 768;   // ds(query i, key x)
    float* prow = lse2 + 1024;   // p(query x, key r)
    float* dsrow = lse2 + 1280;  // ds(query x, key r)
    float* qx = reinterpret_cast<float*>(sm + kBwdOffX);
    float* kx = qx + 64;
    float* vx = qx + 128;
    float* dox = qx + 192;
    float* scal = qx + 256;  // [0] p_xx, [1] ds_xx, [2] lse2_x, [3] delta_x
    // mbarriers: 0 first block's operands landed, 1 all operands landed, 2 S^T / dP^T ready, 3 / 4 first / second half
    // of every warp's P^T and dS^T columns stored, 5 dV_j / dK_j complete, 6 edge row vectors (prow, dsrow) ready,
    // 7 every dQ MMA retired, 8 edge column vectors (pcol, dscol) ready
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kBwdOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, n = blockIdx.y;
    const int D = p.heads * kHd, nv = p.nv, nt = p.n_tiles;
    const size_t vbase = (static_cast<size_t>(n) * p.heads + h) * p.T;
    const size_t cta_id = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
    if (warp == 0) PCG_TRACE(0);

    // One CTA per SM and equal work per CTA: left alone, all 148 CTAs load their 192 KB of operands at the same
    // moment, compute at the same moment and store at the same moment, so the operand loads run at 1/148 of the L2
    // bandwidth each while the L2 idles the rest of the time.  Delaying the first wave by a per-SM offset spreads the
    // phases out; the offsets persist because every SM runs its CTAs back to back.
    if (p.stagger_cycles > 0 && cta_id < static_cast<size_t>(p.wave)) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const long long wait = static_cast<long long>((smid >> 1) & 7) * p.stagger_cycles;
        const long long t0 = clock64();
        while (clock64() - t0 < wait) {
        }
    }

    // delta = rowsum(dO * O) per query needs O, which no MMA reads: the elementwise warps fetch their rows of O and
    // dO straight from global / L2 (eight lanes share a 128-byte row, four rows per step).  Like the operand tiles,
    // the rows are requested in the order they are needed: query block 0 at kernel entry (the latency hides behind
    // the setup), query block 1 once those have been reduced (it is needed after the first block's arithmetic).
    uint4 xo[4], xd[4];
    auto load_o_rows = [&](int half) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int row = half * 128 + warp * 16 + it * 4 + (lane >> 3);
            const size_t off = (static_cast<size_t>(n) * p.T + min(row, nv)) * D + h * kHd + (lane & 7) * 8;
            xo[it] = *reinterpret_cast<const uint4*>(p.out + off);
            xd[it] = *reinterpret_cast<const uint4*>(p.d_out + off);
        }
    };
    if (warp < 8) load_o_rows(0);

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            mbar_init(&bars[2], 1);
            mbar_init(&bars[3], 8);
            mbar_init(&bars[4], 8);
            mbar_init(&bars[5], 1);
            mbar_init(&bars[6], 3);
            mbar_init(&bars[7], 1);
            mbar_init(&bars[8], 3);
            fence_barrier_init();
            mbar_arrive_expect_tx(&bars[0], 4 * kBlkBytes);
            tma_load_3d(&map_qkv, &bars[0], sm_k, D + h * kHd, 0, n, kEvictFirst);
            tma_load_3d(&map_qkv, &bars[0], sm_q, h * kHd, 0, n, kEvictFirst);
            tma_load_3d(&map_qkv, &bars[0], sm_v, 2 * D + h * kHd, 0, n, kEvictFirst);
            tma_load_3d(&map_do, &bars[0], sm_do, h * kHd, 0, n, kEvictFirst);
            // the second tile's operands are requested once the first tile's have landed (below): asking for all
            // 128 KB at once only delays the 64 KB the first block is waiting for
            if (nt == 1) mbar_arrive(&bars[1]);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    } else if (warp < 8) {
        const int t = threadIdx.x;
        lse2[t] = (t < nv) ? p.lse[vbase + t] * kLog2e : INFINITY;
    } else {
        const size_t rowx = static_cast<size_t>(n) * p.T + nv;
        const bf16* xq = p.qkv + rowx * 3 * D + h * kHd;
        if (warp == 9) {
            load_row_f32(qx, xq, lane);
            load_row_f32(dox, p.d_out + rowx * D + h * kHd, lane);
            const uint32_t o2 = *reinterpret_cast<const uint32_t*>(p.out + rowx * D + h * kHd + 2 * lane);
            const float dx = warp_sum(fmaf(bf_lo(o2), dox[2 * lane], bf_hi(o2) * dox[2 * lane + 1]));
            if (lane == 0) scal[2] = p.lse[vbase + nv] * kLog2e, scal[3] = dx;
        }
        if (warp == 10) load_row_f32(kx, xq + D, lane);
        if (warp == 11) load_row_f32(vx, xq + 2 * D, lane);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            const int n_blocks = nt * nt;
            // block b = (key tile j, query block i), j-major; widths of the query blocks
            auto width = [&](int i) { return min(128, ((nv - 128 * i) + 15) & ~15); };
            // pull the operands of the head that runs on this SM next (one wave ahead) into L2
            auto prefetch_next = [&]() {
                const size_t next = cta_id + p.wave;
                if (next < static_cast<size_t>(gridDim.x) * gridDim.y) {
                    const int h2 = static_cast<int>(next % gridDim.x), n2 = static_cast<int>(next / gridDim.x);
                    for (int i = 0; i < nt; ++i) {
                        tma_prefetch_3d(&map_qkv, D + h2 * kHd, i * 128, n2);
                        tma_prefetch_3d(&map_qkv, h2 * kHd, i * 128, n2);
                        tma_prefetch_3d(&map_qkv, 2 * D + h2 * kHd, i * 128, n2);
                        tma_prefetch_3d(&map_do, h2 * kHd, i * 128, n2);
                    }
                }
            };
            mbar_wait(&bars[0], 0);
            PCG_TRACE(1);
            tc_fence_after();
            mma_tile_x_rows(tmem + kColST, sm_k, sm_q, width(0));    // S^T  = K_0 Q_0^T
            mma_tile_x_rows(tmem + kColDPT, sm_v, sm_do, width(0));  // dP^T = V_0 dO_0^T
            umma_commit(&bars[2]);
            PCG_TRACE(21);
            if (nt > 1) {
                mbar_arrive_expect_tx(&bars[1], 4 * kBlkBytes);
                tma_load_3d(&map_qkv, &bars[1], sm_q + kBlkBytes, h * kHd, 128, n, kEvictFirst);
                tma_load_3d(&map_do, &bars[1], sm_do + kBlkBytes, h * kHd, 128, n, kEvictFirst);
                tma_load_3d(&map_qkv, &bars[1], sm_k + kBlkBytes, D + h * kHd, 128, n, kEvictFirst);
                tma_load_3d(&map_qkv, &bars[1], sm_v + kBlkBytes, 2 * D + h * kHd, 128, n, kEvictFirst);
            }
            prefetch_next();  // (behind this head's own loads: 276 -> 269 us per layer)
            mbar_wait(&bars[1], 0);
            tc_fence_after();
            // MMA order per block b: dV / dK on the K-steps whose P^T / dS^T columns are stored (first halves, then second
            // halves) -> S^T, dP^T of block b + 1 -> dQ of block b.  The next block's scores are queued AHEAD of dQ(b) so
            // that the elementwise warps start on them while dQ(b), which reads the other dS^T buffer, still runs; the
            // tensor pipe retires in order, so whatever read the P^T columns or the dS^T buffer about to be refilled
            // is done by the time S^T(b + 1) signals.
            const uint32_t idesc_kn = umma_idesc_bf16(128, 64, 0, 1);
            for (int b = 0; b < n_blocks; ++b) {
                const int j = b / nt, i = b - j * nt;
                const int ksteps = width(i) >> 4;
                const uint8_t* do_i = sm_do + i * kBlkBytes;
                const uint8_t* q_i = sm_q + i * kBlkBytes;
                const uint8_t* dst_b = sm_dst + (b & 1) * 2 * kBlkBytes;
                bool fresh = (i == 0);  // first K-step of a key tile overwrites the accumulators
                // K-step ks = 16 queries: columns [16 ks, +16) of the block; the warps with phase = ks >> 2 store them
                auto grad_step = [&](int ks) {
                    if (ks >= ksteps) return;
                    const uint64_t db_do = umma_smem_desc_sw128(smem_u32(do_i + ks * 2048));
                    const uint64_t db_q = umma_smem_desc_sw128(smem_u32(q_i + ks * 2048));
                    const uint64_t da_ds = umma_smem_desc_sw128(smem_u32(dst_b + (ks >> 2) * kBlkBytes + (ks & 3) * 32));
                    umma_f16_ts(tmem + kColDV, tmem + kColST + 64 * (ks >> 2) + 8 * (ks & 3), db_do, idesc_kn, !fresh);
                    umma_f16(tmem + kColDK, da_ds, db_q, idesc_kn, !fresh);
                    fresh = false;
                };
                mbar_wait(&bars[3], b & 1);
                tc_fence_after();
                grad_step(0), grad_step(4), grad_step(1), grad_step(5);
                // everything stored (and S^T / dP^T consumed)
                mbar_wait(&bars[4], b & 1);
                PCG_TRACE(16 + b);
                tc_fence_after();
                grad_step(2), grad_step(6), grad_step(3), grad_step(7);
                if (i == nt - 1) umma_commit(&bars[5]);  // dV_j, dK_j complete
                if (b + 1 < n_blocks) {
                    const int j2 = (b + 1) / nt, i2 = (b + 1) - j2 * nt;
                    mma_tile_x_rows(tmem + kColST, sm_k + j2 * kBlkBytes, sm_q + i2 * kBlkBytes, width(i2));
                    mma_tile_x_rows(tmem + kColDPT, sm_v + j2 * kBlkBytes, sm_do + i2 * kBlkBytes, width(i2));
                    umma_commit(&bars[2]);
                }
                mma_rows_t_x_cols(tmem + kColDQ + 64 * i, dst_b, sm_k + j * kBlkBytes, 8, j != 0);  // dQ_i += dS K_j
            }
            umma_commit(&bars[7]);
        }
    } else if (warp >= 9) {
        // edge token x = nv: row x (query x against every key) and column x (key x against every query) of the score
        // matrix, one dot product per token and side, then the three matrix-vector products for row x of dV, dK and
        // dQ.  Runs beside the tensor-core pipeline.  Row x comes first: the first key tile's epilogue needs it and it
        // does not depend on delta[]; column x follows once delta[] is complete and is needed by the dQ epilogue.
        mbar_wait(&bars[0], 0);
        mbar_wait(&bars[1], 0);
        const float lse2_x = scal[2], delta_x = scal[3];
        for (int t = (warp - 9) * 32 + lane; t < 256; t += 96) {
            float pr = 0.f, dsr = 0.f;
            if (t < nv) {
                pr = exp2f(fmaf(row_dot(sm_k, t, qx), kLog2e, -lse2_x));
                dsr = pr * (row_dot(sm_v, t, dox) - delta_x);
            }
            prow[t] = pr, dsrow[t] = dsr;
        }
        if (warp == 9 && lane == 0) {
            float sxx = 0.f, dpxx = 0.f;
            for (int d = 0; d < 64; ++d) sxx = fmaf(qx[d], kx[d], sxx), dpxx = fmaf(dox[d], vx[d], dpxx);
            const float pxx = exp2f(fmaf(sxx, kLog2e, -lse2_x));
            scal[0] = pxx;
            scal[1] = pxx * (dpxx - delta_x);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[6]);
        named_bar_sync(1, kBwdThreads - 32);  // delta[] is complete
        for (int t = (warp - 9) * 32 + lane; t < 256; t += 96) {
            float pc = 0.f, dsc = 0.f;
            if (t < nv) {
                pc = exp2f(fmaf(row_dot(sm_q, t, kx), kLog2e, -lse2[t]));
                dsc = pc * (row_dot(sm_do, t, vx) - delta[t]);
            }
            pcol[t] = pc, dscol[t] = dsc;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[8]);
        mbar_wait(&bars[6], 0);
        mbar_wait(&bars[8], 0);
        if (warp == 9) PCG_TRACE(2);
        bf16* gx = p.d_qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + h * kHd;
        if (warp == 9) edge_gemv(pcol, sm_do, nv, scal[0], dox, gx + 2 * D, 1.0f, lane);  // dV_x
        if (warp == 10) edge_gemv(dscol, sm_q, nv, scal[1], qx, gx + D, 1.0f, lane);       // dK_x
        if (warp == 11) edge_gemv(dsrow, sm_k, nv, scal[1], kx, gx, 1.0f, lane);           // dQ_x
        if (warp == 11) PCG_TRACE(3);
    } else {
        const int quarter = warp & 3, phase = warp >> 2;
        const int r = quarter * 32 + lane;  // key row in the tile == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        uint8_t* stage = sm + kBwdOffStage + warp * 2048;
        bf16* gd = p.d_qkv + static_cast<size_t>(n) * p.T * 3 * D + h * kHd + phase * 32;
        // delta = rowsum(dO * O): reduce the loaded chunks over the 8 lanes of each row
        auto reduce_delta = [&](int half) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int row = half * 128 + warp * 16 + it * 4 + (lane >> 3);
                float d = chunk_dot(xo[it], xd[it]);
                d += __shfl_xor_sync(0xffffffffu, d, 1);
                d += __shfl_xor_sync(0xffffffffu, d, 2);
                d += __shfl_xor_sync(0xffffffffu, d, 4);
                if ((lane & 7) == 0) delta[row] = (row < nv) ? d : 0.f;
            }
        };
        reduce_delta(0);
        load_o_rows(1);
        {
            const size_t next = cta_id + p.wave;  // the next head's O rows (its dO rows come with the TMA prefetch)
            if (next < static_cast<size_t>(gridDim.x) * gridDim.y && threadIdx.x <= nv)
                prefetch_l2(p.out + (static_cast<size_t>(next / gridDim.x) * p.T + threadIdx.x) * D + (next % gridDim.x) * kHd);
        }
        named_bar_sync(2, 256);  // delta of query block 0 is complete (the elementwise warps only)
        if (warp == 0) PCG_TRACE(20);
        int b = 0;
        for (int j = 0; j < nt; ++j) {
            const bool row_ok = (j * 128 + r) < nv;
            for (int i = 0; i < nt; ++i, ++b) {
                const int width = min(128, ((nv - 128 * i) + 15) & ~15);
                const float* l2 = lse2 + i * 128;
                const float* dl = delta + i * 128;
                mbar_wait(&bars[2], b & 1);
                tc_fence_after();
                if (warp == 0) PCG_TRACE(4 + 2 * b);
                // this warp's columns: the 64-column block `phase`, 16 at a time with the next TMEM load in flight behind
                // the arithmetic.  P^T goes back into tensor memory over S^T columns this warp has already read
                // (packed: 8 columns per 16 queries), dS^T into this block's shared-memory buffer.
                const int cb = 64 * phase;
                uint8_t* dst_blk = sm_dst + (b & 1) * 2 * kBlkBytes + phase * kBlkBytes;
                uint32_t sa[16], da[16], sb[16], db[16];
                auto finish = [&](const uint32_t(&sv)[16], const uint32_t(&dv)[16], int g) {
                    const Cols16 o = bwd_cols16(sv, dv, l2 + cb + 16 * g, dl + cb + 16 * g, row_ok);
                    bwd_store_ds(o, dst_blk, r, 2 * g);
                    const uint32_t pk[8] = {o.p[0].x, o.p[0].y, o.p[0].z, o.p[0].w, o.p[1].x, o.p[1].y, o.p[1].z, o.p[1].w};
                    tmem_st<8>(trow + kColST + cb + 8 * g, pk);
                };
                if (cb < width) tmem_ld<16>(trow + kColST + cb, sa), tmem_ld<16>(trow + kColDPT + cb, da);
                tmem_wait_ld();
                if (cb + 16 < width) tmem_ld<16>(trow + kColST + cb + 16, sb), tmem_ld<16>(trow + kColDPT + cb + 16, db);
                if (cb < width) finish(sa, da, 0);
                tmem_wait_ld();
                if (cb + 32 < width) tmem_ld<16>(trow + kColST + cb + 32, sa), tmem_ld<16>(trow + kColDPT + cb + 32, da);
                if (cb + 16 < width) finish(sb, db, 1);
                tmem_wait_st();
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[3]);
                tmem_wait_ld();
                if (cb + 48 < width) tmem_ld<16>(trow + kColST + cb + 48, sb), tmem_ld<16>(trow + kColDPT + cb + 48, db);
                if (cb + 32 < width) finish(sa, da, 2);
                tmem_wait_ld();
                if (cb + 48 < width) finish(sb, db, 3);
                tmem_wait_st();
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[4]);
                if (warp == 0) PCG_TRACE(5 + 2 * b);
                if (b == 0) {
                    reduce_delta(1);
                    named_bar_sync(1, kBwdThreads - 32);  // all of delta[] is complete (with the edge warps)
                }
            }
            // key tile j finished: dV_j and dK_j (+ the edge query's contribution) -> global
            if (j == 0) mbar_wait(&bars[6], 0);
            mbar_wait(&bars[5], j & 1);
            tc_fence_after();
            if (warp == 0) PCG_TRACE(12 + j);
            const int row0 = j * 128 + quarter * 32;
            bwd_epilogue(trow + kColDV + phase * 32, prow[j * 128 + r], dox + phase * 32, stage, lane, gd + 2 * D,
                         static_cast<size_t>(3) * D, row0, nv);
            bwd_epilogue(trow + kColDK + phase * 32, dsrow[j * 128 + r], qx + phase * 32, stage, lane, gd + D,
                         static_cast<size_t>(3) * D, row0, nv);
            tc_fence_before();
        }
        // dQ_i (+ the edge key's contribution)
        mbar_wait(&bars[8], 0);
        mbar_wait(&bars[7], 0);
        tc_fence_after();
        for (int i = 0; i < nt; ++i)
            bwd_epilogue(trow + kColDQ + 64 * i + phase * 32, dscol[i * 128 + r], kx + phase * 32, stage, lane, gd,
                         static_cast<size_t>(3) * D, i * 128 + quarter * 32, nv);
        if (warp == 0) PCG_TRACE(14);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) PCG_TRACE(15);
    if (warp == 8) tmem_dealloc(tmem, 512);
}

// Two independent 32-column accumulator slices at once (the persistent backward's drains): both tensor-memory loads
// are issued before either is consumed, and the rows go straight to global memory from the tensor-memory layout
// (thread = row): a drain is one warp's dependent chain, and the staging round trip that makes the one-shot kernel's
// stores row-coalesced (store -> sync -> load -> store, ~600 clk) costs more here than the extra store transactions.
// Out of line: the three drains share one copy of the code.
__device__ __noinline__ void bwd_epilogue2(uint32_t taddr_a, float coef_a, const float* xrow_a, bf16* gbase_a, int row0_a,
                                           uint32_t taddr_b, float coef_b, const float* xrow_b, bf16* gbase_b, int row0_b,
                                           int lane, size_t ld, int row_end) {
    uint32_t va[32], vb[32];
    tmem_ld<32>(taddr_a, va);
    tmem_ld<32>(taddr_b, vb);
    tmem_wait_ld();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t(&v)[32] = half == 0 ? va : vb;
        const float coef = half == 0 ? coef_a : coef_b;
        const float* xrow = half == 0 ? xrow_a : xrow_b;
        const int row = (half == 0 ? row0_a : row0_b) + lane;
        bf16* grow = (half == 0 ? gbase_a : gbase_b) + static_cast<size_t>(row) * ld;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 xa = *reinterpret_cast<const float4*>(xrow + g * 8);
            const float4 xb = *reinterpret_cast<const float4*>(xrow + g * 8 + 4);
            const uint4 o = make_uint4(pack_bf16(fmaf(coef, xa.x, __uint_as_float(v[8 * g])),
                                                 fmaf(coef, xa.y, __uint_as_float(v[8 * g + 1]))),
                                       pack_bf16(fmaf(coef, xa.z, __uint_as_float(v[8 * g + 2])),
                                                 fmaf(coef, xa.w, __uint_as_float(v[8 * g + 3]))),
                                       pack_bf16(fmaf(coef, xb.x, __uint_as_float(v[8 * g + 4])),
                                                 fmaf(coef, xb.y, __uint_as_float(v[8 * g + 5]))),
                                       pack_bf16(fmaf(coef, xb.z, __uint_as_float(v[8 * g + 6])),
                                                 fmaf(coef, xb.w, __uint_as_float(v[8 * g + 7]))));
            // thread = row: four 16-byte stores cover this row's 64 contiguous bytes (two full 32-byte sectors)
            if (row < row_end) *reinterpret_cast<uint4*>(grow + g * 8) = o;
        }
    }
}

// One 64-column accumulator (dV_j, dK_j or dQ_i) of this warp's 32 rows -> bf16, through a 4 KB swizzled staging tile and
// ONE bulk tensor store.  Storing rows straight from the tensor-memory layout (bwd_epilogue2 above: thread = row, 16
// bytes per lane and instruction) costs 32 L1 wavefronts per store instruction -- 6144 per item, ~12 000 clk of the
// load/store unit at the measured 2 clk per wavefront, which the elementwise and edge warps queue behind; the staging
// tile takes 4 wavefronts per instruction and the TMA engine writes whole lines.  Rows >= nv are clipped by the map.
__device__ __noinline__ void bwd_drain_tma(uint32_t taddr, float coef, const float* xrow, uint8_t* stage, const CUtensorMap* map,
                                           int col, int row0, int n, int lane) {
    uint32_t va[32], vb[32];
    tmem_ld<32>(taddr, va);
    tmem_ld<32>(taddr + 32, vb);
    tmem_wait_ld();
    if (lane == 0) bulk_wait_read_all();  // the previous store of this warp has read the staging tile
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t(&v)[32] = half == 0 ? va : vb;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 xa = *reinterpret_cast<const float4*>(xrow + half * 32 + g * 8);
            const float4 xb = *reinterpret_cast<const float4*>(xrow + half * 32 + g * 8 + 4);
            const uint4 o = make_uint4(pack_bf16(fmaf(coef, xa.x, __uint_as_float(v[8 * g])),
                                                 fmaf(coef, xa.y, __uint_as_float(v[8 * g + 1]))),
                                       pack_bf16(fmaf(coef, xa.z, __uint_as_float(v[8 * g + 2])),
                                                 fmaf(coef, xa.w, __uint_as_float(v[8 * g + 3]))),
                                       pack_bf16(fmaf(coef, xb.x, __uint_as_float(v[8 * g + 4])),
                                                 fmaf(coef, xb.y, __uint_as_float(v[8 * g + 5]))),
                                       pack_bf16(fmaf(coef, xb.z, __uint_as_float(v[8 * g + 6])),
                                                 fmaf(coef, xb.w, __uint_as_float(v[8 * g + 7]))));
            *reinterpret_cast<uint4*>(stage + row_chunk(lane, half * 4 + g)) = o;
        }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        tma_store_3d(map, stage, col, row0, n);
        bulk_commit_group();
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward, persistent (the production kernel for 130 <= T <= 257: two key tiles, two query blocks)
// ---------------------------------------------------------------------------------------------------------
// Same arithmetic, tensor-memory plan and block order as attn_bwd_tc_kernel, but one CTA per SM walks the (cutout,
// head) items w = blockIdx.x + it * gridDim.x, and three things that kernel does in sequence run beside the block
// pipeline here (measured on its phase trace: ~5000 clk of operand loads + delta before the first block, ~2800 clk of
// dV_0 / dK_0 epilogue between the key tiles, ~4200 clk of epilogues at the end, of a 28 300 clk CTA):
//   * the operand tiles of item it + 1 are requested as soon as the last MMA that reads their slot has retired
//     (V_0 during block 1, K_0 during block 2, Q_0 / dO_0 during block 3, the second halves at the item's end), so the
//     first scores of the next item are issued right behind the last block's gradient products;
//   * four auxiliary warps (one per tensor-memory lane quarter) own everything that is not the 128 x 128 block
//     arithmetic: the edge token's row and column, lse / delta / edge vectors of the NEXT item (double buffered), and
//     all six accumulator epilogues -- the eight elementwise warps go from block to block without ever draining;
//   * barriers complete once per block (S ready, P / dS halves stored), per key tile (dV_j, dK_j complete / drained)
//     or per item, and every wait names the completion it needs by its running index.
// Warps: 0-7 elementwise (lane quarter = warp & 3, column phase = warp >> 2), 8 TMA + MMA issue, 9-12 auxiliary
// (lane quarter = warp & 3 = 1, 2, 3, 0).
constexpr int kPbEdge = 3;   // edge warps 9..11: the edge token's row and column, next item's lse / edge vectors
constexpr int kPbDrain = 4;  // drain warps 12..15 (lane quarters 0..3): accumulator epilogues, next item's delta
constexpr int kPbAux = kPbEdge + kPbDrain;  // 16 warps = 4 per scheduler: 128 registers (a 17th caps the kernel at 96)
constexpr int kPbThreads = (9 + kPbAux) * 32;
constexpr int kPbOffVec = 12 * kBlkBytes;                 // 2 slots x float[256] x 6: lse2, delta, pcol, dscol, prow, dsrow
constexpr int kPbOffX = kPbOffVec + 2 * 6 * 1024;         // 2 slots x {float[64] x 4: q_x, k_x, v_x, dO_x; 8 scalars}
constexpr int kPbXFloats = 4 * 64 + 8;
constexpr int kPbOffBar = kPbOffX + 2 * kPbXFloats * 4;
constexpr int kPbOffStage = (kPbOffBar + 20 * 8 + 64 + 1023) / 1024 * 1024;  // <= 20 mbarriers + tmem slot; then 4 x 4 KB
constexpr int kPbSmemBytes = kPbOffStage + kPbDrain * 4096 + 1024;           // drain staging tiles (128-byte swizzle)
static_assert(kPbSmemBytes <= 232448, "persistent backward exceeds the 227 KB dynamic shared memory limit");
enum {
    kPbFullK0 = 0, kPbFullV0 = 1, kPbFullQD0 = 2, kPbFull1 = 3,  // operand tiles landed (TMA)
    kPbSReady = 4,      // S^T / dP^T of a block retired                       (per block)
    kPbHalf1 = 5,       // first half of every warp's P^T / dS^T columns stored (per block, 8 warps)
    kPbHalf2 = 6,       // second half                                          (per block, 8 warps)
    kPbTileDone = 7,    // dV_j / dK_j complete                                 (per key tile)
    kPbDqDone = 8,      // every MMA of the item retired                        (per item)
    kPbAccFreeVK = 9,   // dV / dK drained by the 4 aux warps                   (per key tile)
    kPbAccFreeQ = 10,   // dQ_0 / dQ_1 drained                                  (per item)
    kPbEdgeDone = 11,   // aux warps are done reading the operand tiles         (per item)
    kPbVecReady = 12,   // [2] lse2 / delta / edge vectors of the slot's item   (per use of the slot)
    kPbGateK0 = 14,     // last reader of K_0 retired                           (per item)
    kPbGateQD0 = 15,    // last reader of Q_0 / dO_0 retired                    (per item)
    kPbSlotFree = 16,   // [2] drain warps are done with the slot's vectors     (per use of the slot)
    kPbRowA = 18,       // edge row of key tile 0 (prow / dsrow, t < 128) ready  (per item, 3 edge warps)
    kPbCount = 19
};

struct PbParams {
    int T, heads;
    int nv;     // T - 1 in (128, 256]
    int items;  // n * heads
    const bf16* qkv;
    const bf16* out;
    const bf16* d_out;
    const float* lse;
    bf16* d_qkv;
    long long* trace;
};

__global__ void __launch_bounds__(kPbThreads, 1)
attn_bwd_persist_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                        const __grid_constant__ CUtensorMap map_dqkv, const PbParams p) {
    grid_dep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_q = sm + kBwdOffQ;
    uint8_t* sm_k = sm + kBwdOffK;
    uint8_t* sm_v = sm + kBwdOffV;
    uint8_t* sm_do = sm + kBwdOffDO;
    uint8_t* sm_dst = sm + kBwdOffDST;
    float* vec0 = reinterpret_cast<float*>(sm + kPbOffVec);  // slot s: + 1536 s; lse2, delta, pcol, dscol, prow, dsrow
    float* x0 = reinterpret_cast<float*>(sm + kPbOffX);       // slot s: + kPbXFloats s; q_x, k_x, v_x, dO_x, scalars
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kPbOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kPbCount);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.heads * kHd, nv = p.nv;
    const int G = gridDim.x;
    const int n_my = (p.items - static_cast<int>(blockIdx.x) + G - 1) / G;
    const size_t cta_id = blockIdx.x;
    const int w1 = min(128, ((nv - 128) + 15) & ~15);  // width of query block 1 (block 0 is full)
    auto decode = [&](int it, int& n, int& h) {
        const int w = static_cast<int>(blockIdx.x) + it * G;
        n = w / p.heads;
        h = w - n * p.heads;
    };

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(&bars[kPbFullK0], 1);
            mbar_init(&bars[kPbFullV0], 1);
            mbar_init(&bars[kPbFullQD0], 1);
            mbar_init(&bars[kPbFull1], 1);
            mbar_init(&bars[kPbSReady], 1);
            mbar_init(&bars[kPbHalf1], 8);
            mbar_init(&bars[kPbHalf2], 8);
            mbar_init(&bars[kPbTileDone], 1);
            mbar_init(&bars[kPbDqDone], 1);
            mbar_init(&bars[kPbAccFreeVK], kPbDrain);
            mbar_init(&bars[kPbAccFreeQ], kPbDrain);
            mbar_init(&bars[kPbEdgeDone], kPbEdge);
            mbar_init(&bars[kPbVecReady], kPbAux);
            mbar_init(&bars[kPbVecReady + 1], kPbAux);
            mbar_init(&bars[kPbSlotFree], kPbDrain);
            mbar_init(&bars[kPbSlotFree + 1], kPbDrain);
            mbar_init(&bars[kPbRowA], kPbEdge);
            mbar_init(&bars[kPbGateK0], 1);
            mbar_init(&bars[kPbGateQD0], 1);
            fence_barrier_init();
            int n, h;
            decode(0, n, h);
            mbar_arrive_expect_tx(&bars[kPbFullK0], kBlkBytes);
            tma_load_3d(&map_qkv, &bars[kPbFullK0], sm_k, D + h * kHd, 0, n, kEvictFirst);
            mbar_arrive_expect_tx(&bars[kPbFullQD0], 2 * kBlkBytes);
            tma_load_3d(&map_qkv, &bars[kPbFullQD0], sm_q, h * kHd, 0, n, kEvictFirst);
            tma_load_3d(&map_do, &bars[kPbFullQD0], sm_do, h * kHd, 0, n, kEvictFirst);
            mbar_arrive_expect_tx(&bars[kPbFullV0], kBlkBytes);
            tma_load_3d(&map_qkv, &bars[kPbFullV0], sm_v, 2 * D + h * kHd, 0, n, kEvictFirst);
            mbar_arrive_expect_tx(&bars[kPbFull1], 4 * kBlkBytes);
            tma_load_3d(&map_qkv, &bars[kPbFull1], sm_q + kBlkBytes, h * kHd, 128, n, kEvictFirst);
            tma_load_3d(&map_do, &bars[kPbFull1], sm_do + kBlkBytes, h * kHd, 128, n, kEvictFirst);
            tma_load_3d(&map_qkv, &bars[kPbFull1], sm_k + kBlkBytes, D + h * kHd, 128, n, kEvictFirst);
            tma_load_3d(&map_qkv, &bars[kPbFull1], sm_v + kBlkBytes, 2 * D + h * kHd, 128, n, kEvictFirst);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[cta_id * 32 + 30] = clock64(), p.trace[cta_id * 32 + 29] = n_my;

    if (warp == 8) {
        // ================================================================ TMA + MMA issue (one thread)
        if (lane == 0) {
            const uint32_t idesc_kn = umma_idesc_bf16(128, 64, 0, 1);
            auto width = [&](int i) { return i == 0 ? 128 : w1; };
            auto scores = [&](int j, int i) {  // S^T = K_j Q_i^T, dP^T = V_j dO_i^T
                mma_tile_x_rows(tmem + kColST, sm_k + j * kBlkBytes, sm_q + i * kBlkBytes, width(i));
                mma_tile_x_rows(tmem + kColDPT, sm_v + j * kBlkBytes, sm_do + i * kBlkBytes, width(i));
                umma_commit(&bars[kPbSReady]);
            };
#pragma unroll 1
            for (int it = 0; it < n_my; ++it) {
                const uint32_t ip = it & 1;
                const bool has_next = it + 1 < n_my;
                int n2 = 0, h2 = 0;
                if (has_next) decode(it + 1, n2, h2);
                if (it == 0) {  // later items: issued behind the previous item's last gradient products (below)
                    mbar_wait_c(&bars[kPbFullK0], 0);
                    mbar_wait_c(&bars[kPbFullQD0], 0);
                    mbar_wait_c(&bars[kPbFullV0], 0);
                    tc_fence_after();
                    scores(0, 0);
                }
                bool next_scores_issued = false;
#pragma unroll 1
                for (int b = 0; b < 4; ++b) {
                    const int g = it * 4 + b, j = b >> 1, i = b & 1;
                    const int ksteps = width(i) >> 4;
                    const uint8_t* do_i = sm_do + i * kBlkBytes;
                    const uint8_t* q_i = sm_q + i * kBlkBytes;
                    const uint8_t* dst_b = sm_dst + (b & 1) * 2 * kBlkBytes;
                    bool fresh = (i == 0);  // first K-step of a key tile overwrites the accumulators
                    auto grad_step = [&](int ks) {  // K-step ks = 16 queries: columns [16 ks, +16) of the block
                        if (ks >= ksteps) return;
                        const uint64_t db_do = umma_smem_desc_sw128(smem_u32(do_i + ks * 2048));
                        const uint64_t db_q = umma_smem_desc_sw128(smem_u32(q_i + ks * 2048));
                        const uint64_t da_ds = umma_smem_desc_sw128(smem_u32(dst_b + (ks >> 2) * kBlkBytes + (ks & 3) * 32));
                        umma_f16_ts(tmem + kColDV, tmem + kColST + 64 * (ks >> 2) + 8 * (ks & 3), db_do, idesc_kn, !fresh);
                        umma_f16(tmem + kColDK, da_ds, db_q, idesc_kn, !fresh);
                        fresh = false;
                    };
                    mbar_wait_c(&bars[kPbHalf1], g & 1);
                    if (it == kTraceItem && p.trace != nullptr) p.trace[cta_id * 32 + 20 + b] = clock64();
                    if (i == 0) {  // dV / dK are about to be overwritten: drain number it * 2 + j - 1 must be complete
                        const int d = it * 2 + j;
                        if (d > 0) mbar_wait_c(&bars[kPbAccFreeVK], (d - 1) & 1);
                    }
                    // operand tiles of the next item whose slots are free by now
                    if (has_next) {
                        if (b == 3) {
                            // K_0 / V_0 / Q_0 / dO_0 have no reader left in this item: the MMAs that read them have retired
                            // (gates) and the edge warps have taken the edge token's dot products from every tile
                            mbar_wait_c(&bars[kPbEdgeDone], ip);
                            mbar_wait_c(&bars[kPbGateK0], ip);
                            mbar_wait_c(&bars[kPbGateQD0], ip);
                            mbar_arrive_expect_tx(&bars[kPbFullK0], kBlkBytes);
                            tma_load_3d(&map_qkv, &bars[kPbFullK0], sm_k, D + h2 * kHd, 0, n2, kEvictFirst);
                            mbar_arrive_expect_tx(&bars[kPbFullQD0], 2 * kBlkBytes);
                            tma_load_3d(&map_qkv, &bars[kPbFullQD0], sm_q, h2 * kHd, 0, n2, kEvictFirst);
                            tma_load_3d(&map_do, &bars[kPbFullQD0], sm_do, h2 * kHd, 0, n2, kEvictFirst);
                            mbar_arrive_expect_tx(&bars[kPbFullV0], kBlkBytes);
                            tma_load_3d(&map_qkv, &bars[kPbFullV0], sm_v, 2 * D + h2 * kHd, 0, n2, kEvictFirst);
                        }
                    }
                    tc_fence_after();
                    if (it == kTraceItem && p.trace != nullptr) p.trace[cta_id * 32 + 24 + b] = clock64();
                    grad_step(0), grad_step(4), grad_step(1), grad_step(5);
                    mbar_wait_c(&bars[kPbHalf2], g & 1);  // everything stored (and S^T / dP^T consumed)
                    tc_fence_after();
                    grad_step(2), grad_step(6), grad_step(3), grad_step(7);
                    if (b == 2) umma_commit(&bars[kPbGateQD0]);  // Q_0 / dO_0 have no reader left in this item
                    if (i == 1) umma_commit(&bars[kPbTileDone]);  // dV_j, dK_j complete
                    // the next block's scores go AHEAD of dQ(b): the elementwise warps start on them while dQ(b) runs
                    if (b == 0) {
                        mbar_wait_c(&bars[kPbFull1], ip);  // Q_1 / dO_1 (and K_1 / V_1)
                        tc_fence_after();
                        scores(0, 1);
                    } else if (b == 1) {
                        scores(1, 0);
                    } else if (b == 2) {
                        scores(1, 1);
                    } else if (has_next && mbar_test_wait(&bars[kPbFullK0], ip ^ 1) && mbar_test_wait(&bars[kPbFullQD0], ip ^ 1) &&
                               mbar_test_wait(&bars[kPbFullV0], ip ^ 1)) {
                        tc_fence_after();
                        scores(0, 0);  // first block of the next item (its operands sit in the slots freed above)
                        next_scores_issued = true;
                    }
                    if (b == 0 && it > 0) mbar_wait_c(&bars[kPbAccFreeQ], (it - 1) & 1);  // dQ_0 / dQ_1 of the last item drained
                    mma_rows_t_x_cols(tmem + kColDQ + 64 * i, dst_b, sm_k + j * kBlkBytes, 8, j != 0);  // dQ_i += dS K_j
                    if (b == 1) umma_commit(&bars[kPbGateK0]);  // dQ(b1) was the last reader of K_0
                }
                umma_commit(&bars[kPbDqDone]);
                if (has_next) {
                    if (!next_scores_issued) {
                        mbar_wait_c(&bars[kPbFullK0], ip ^ 1);
                        mbar_wait_c(&bars[kPbFullQD0], ip ^ 1);
                        mbar_wait_c(&bars[kPbFullV0], ip ^ 1);
                        tc_fence_after();
                        scores(0, 0);
                    }
                    // every MMA of this item has retired: the second-half slots take the next item's tiles
                    mbar_wait_c(&bars[kPbDqDone], ip);
                    mbar_arrive_expect_tx(&bars[kPbFull1], 4 * kBlkBytes);
                    tma_load_3d(&map_qkv, &bars[kPbFull1], sm_q + kBlkBytes, h2 * kHd, 128, n2, kEvictFirst);
                    tma_load_3d(&map_do, &bars[kPbFull1], sm_do + kBlkBytes, h2 * kHd, 128, n2, kEvictFirst);
                    tma_load_3d(&map_qkv, &bars[kPbFull1], sm_k + kBlkBytes, D + h2 * kHd, 128, n2, kEvictFirst);
                    tma_load_3d(&map_qkv, &bars[kPbFull1], sm_v + kBlkBytes, 2 * D + h2 * kHd, 128, n2, kEvictFirst);
                }
            }
        }
    } else if (warp < 8) {
        // ================================================================ elementwise warps: P^T and dS^T, block after block
        const int quarter = warp & 3, phase = warp >> 2;
        const int r = quarter * 32 + lane;  // key row in the tile == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        const int cb = 64 * phase;  // this warp's columns of every block
#pragma unroll 1
        for (int it = 0; it < n_my; ++it) {
            const int s = it & 1;
            const float* lse2 = vec0 + 1536 * s;
            const float* delta = lse2 + 256;
            const bool tr = (it == kTraceItem) && warp == 0 && lane == 0 && p.trace != nullptr;
            if (it == kTraceItem + 1 && warp == 0 && lane == 0 && p.trace != nullptr) p.trace[cta_id * 32 + 9] = clock64();
            if (tr) p.trace[cta_id * 32 + 0] = clock64();
            mbar_wait_c(&bars[kPbVecReady + s], (it >> 1) & 1);
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int g = it * 4 + b, j = b >> 1, i = b & 1;
                const bool row_ok = (j * 128 + r) < nv;
                const int width = (i == 0) ? 128 : w1;
                const float* l2 = lse2 + i * 128;
                const float* dl = delta + i * 128;
                mbar_wait_c(&bars[kPbSReady], g & 1);
                tc_fence_after();
                if (tr) p.trace[cta_id * 32 + 1 + 2 * b] = clock64();
                uint8_t* dst_blk = sm_dst + (b & 1) * 2 * kBlkBytes + phase * kBlkBytes;
                uint32_t sa[16], da[16], sb[16], db[16];
                auto finish = [&](const uint32_t(&sv)[16], const uint32_t(&dv)[16], int q) {
                    const Cols16 o = bwd_cols16(sv, dv, l2 + cb + 16 * q, dl + cb + 16 * q, row_ok);
                    bwd_store_ds(o, dst_blk, r, 2 * q);
                    const uint32_t pk[8] = {o.p[0].x, o.p[0].y, o.p[0].z, o.p[0].w, o.p[1].x, o.p[1].y, o.p[1].z, o.p[1].w};
                    tmem_st<8>(trow + kColST + cb + 8 * q, pk);
                };
                if (cb < width) tmem_ld<16>(trow + kColST + cb, sa), tmem_ld<16>(trow + kColDPT + cb, da);
                tmem_wait_ld();
                if (cb + 16 < width) tmem_ld<16>(trow + kColST + cb + 16, sb), tmem_ld<16>(trow + kColDPT + cb + 16, db);
                if (cb < width) finish(sa, da, 0);
                tmem_wait_ld();
                if (cb + 32 < width) tmem_ld<16>(trow + kColST + cb + 32, sa), tmem_ld<16>(trow + kColDPT + cb + 32, da);
                if (cb + 16 < width) finish(sb, db, 1);
                tmem_wait_st();
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[kPbHalf1]);
                tmem_wait_ld();
                if (cb + 48 < width) tmem_ld<16>(trow + kColST + cb + 48, sb), tmem_ld<16>(trow + kColDPT + cb + 48, db);
                if (cb + 32 < width) finish(sa, da, 2);
                tmem_wait_ld();
                if (cb + 48 < width) finish(sb, db, 3);
                tmem_wait_st();
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[kPbHalf2]);
                if (tr) p.trace[cta_id * 32 + 2 + 2 * b] = clock64();
            }
        }
    } else {
        // ================================================================ auxiliary warps: two independent groups
        // pull the rows the preparation of item `it` reads (O, dO, lse, the edge token's q / k / v) into L2 well ahead:
        // under load an HBM round trip is ~3000 clk
        auto prefetch_rows = [&](int it, int first, int stride) {
            int n, h;
            decode(it, n, h);
            for (int t = first; t <= nv; t += stride) {
                const size_t off = (static_cast<size_t>(n) * p.T + t) * D + h * kHd;
                prefetch_l2(p.out + off);
                prefetch_l2(p.d_out + off);
            }
            if (first < 3) prefetch_l2(p.qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + first * D + h * kHd);
            if (first < 9) prefetch_l2(p.lse + (static_cast<size_t>(n) * p.heads + h) * p.T + min(first * 32, nv));
        };
        if (warp < 9 + kPbEdge) {
            // ------------------------------------------------------------ edge warps
            const int ew = warp - 9;         // 0..2
            const int el = ew * 32 + lane;   // 0..95
            // lse2 and the edge token's vectors / scalars of item `it` into its slot (global memory only)
            auto prepare_vectors = [&](int it) {
                const int s = it & 1;
                int n, h;
                decode(it, n, h);
                float* lse2 = vec0 + 1536 * s;
                float* qx = x0 + kPbXFloats * s;
                float* kx = qx + 64;
                float* vx = qx + 128;
                float* dox = qx + 192;
                float* scal = qx + 256;
                const size_t vbase = (static_cast<size_t>(n) * p.heads + h) * p.T;
                const size_t rowx = static_cast<size_t>(n) * p.T + nv;
                const bf16* xq = p.qkv + rowx * 3 * D + h * kHd;
                uint32_t e0 = 0, e1 = 0, e2 = 0;
                float lse_x = 0.f, l[3];
                if (ew == 0) {
                    e0 = *reinterpret_cast<const uint32_t*>(xq + 2 * lane);                              // q_x
                    e1 = *reinterpret_cast<const uint32_t*>(p.d_out + rowx * D + h * kHd + 2 * lane);    // dO_x
                    e2 = *reinterpret_cast<const uint32_t*>(p.out + rowx * D + h * kHd + 2 * lane);      // O_x
                    lse_x = p.lse[vbase + nv];
                } else if (ew == 1) {
                    e0 = *reinterpret_cast<const uint32_t*>(xq + D + 2 * lane);                          // k_x
                } else {
                    e0 = *reinterpret_cast<const uint32_t*>(xq + 2 * D + 2 * lane);                      // v_x
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) l[u] = (el + 96 * u < nv) ? p.lse[vbase + el + 96 * u] : INFINITY;
#pragma unroll
                for (int u = 0; u < 3; ++u)
                    if (el + 96 * u < 256) lse2[el + 96 * u] = l[u] * kLog2e;
                if (ew == 0) {
                    qx[2 * lane] = bf_lo(e0), qx[2 * lane + 1] = bf_hi(e0);
                    dox[2 * lane] = bf_lo(e1), dox[2 * lane + 1] = bf_hi(e1);
                    const float dx = warp_sum(fmaf(bf_lo(e2), bf_lo(e1), bf_hi(e2) * bf_hi(e1)));
                    if (lane == 0) scal[2] = lse_x * kLog2e, scal[3] = dx;
                } else if (ew == 1) {
                    kx[2 * lane] = bf_lo(e0), kx[2 * lane + 1] = bf_hi(e0);
                } else {
                    vx[2 * lane] = bf_lo(e0), vx[2 * lane + 1] = bf_hi(e0);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[kPbVecReady + s]);
            };
            prepare_vectors(0);
#pragma unroll 1
            for (int it = 0; it < n_my; ++it) {
                const int s = it & 1;
                const uint32_t ip = it & 1;
                int n, h;
                decode(it, n, h);
                float* lse2 = vec0 + 1536 * s;
                float* delta = lse2 + 256;
                float* pcol = lse2 + 512;    // p(query i, key x)
                float* dscol = lse2 + 768;   // ds(query i, key x)
                float* prow = lse2 + 1024;   // p(query x, key r)
                float* dsrow = lse2 + 1280;  // ds(query x, key r)
                float* qx = x0 + kPbXFloats * s;
                float* kx = qx + 64;
                float* vx = qx + 128;
                float* dox = qx + 192;
                float* scal = qx + 256;  // [0] p_xx, [1] ds_xx, [2] lse2_x, [3] delta_x
                const bool tr = (it == kTraceItem) && ew == 0 && lane == 0 && p.trace != nullptr;
                if (tr) p.trace[cta_id * 32 + 10] = clock64();
                if (it + 1 < n_my) prefetch_rows(it + 1, el, 96);
                mbar_wait_c(&bars[kPbVecReady + s], (it >> 1) & 1);  // lse2, delta (drain warps), edge vectors
                // edge token x = nv: row x (query x against every key) and column x (key x against every query).  Key tile
                // 0's part of row x comes first and is announced on its own: the drain of dV_0 / dK_0 needs nothing else.
                const float lse2_x = scal[2], delta_x = scal[3];
                mbar_wait_c(&bars[kPbFullK0], ip);
                mbar_wait_c(&bars[kPbFullV0], ip);
                if (tr) p.trace[cta_id * 32 + 11] = clock64();
#pragma unroll 1
                for (int t = el; t < 128; t += 96) {
                    const float pr = exp2f(fmaf(row_dot(sm_k, t, qx), kLog2e, -lse2_x));
                    prow[t] = pr, dsrow[t] = pr * (row_dot(sm_v, t, dox) - delta_x);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[kPbRowA]);
                mbar_wait_c(&bars[kPbFullQD0], ip);
#pragma unroll 1
                for (int t = el; t < 128; t += 96) {
                    const float pc = exp2f(fmaf(row_dot(sm_q, t, kx), kLog2e, -lse2[t]));
                    pcol[t] = pc, dscol[t] = pc * (row_dot(sm_do, t, vx) - delta[t]);
                }
                mbar_wait_c(&bars[kPbFull1], ip);
#pragma unroll 1
                for (int t = 128 + el; t < 256; t += 96) {
                    float pr = 0.f, dsr = 0.f, pc = 0.f, dsc = 0.f;
                    if (t < nv) {
                        pr = exp2f(fmaf(row_dot(sm_k, t, qx), kLog2e, -lse2_x));
                        dsr = pr * (row_dot(sm_v, t, dox) - delta_x);
                        pc = exp2f(fmaf(row_dot(sm_q, t, kx), kLog2e, -lse2[t]));
                        dsc = pc * (row_dot(sm_do, t, vx) - delta[t]);
                    }
                    prow[t] = pr, dsrow[t] = dsr, pcol[t] = pc, dscol[t] = dsc;
                }
                if (ew == 2) {  // q_x . k_x and dO_x . v_x (two elements per lane)
                    const float sxx = warp_sum(fmaf(qx[2 * lane], kx[2 * lane], qx[2 * lane + 1] * kx[2 * lane + 1]));
                    const float dpxx = warp_sum(fmaf(dox[2 * lane], vx[2 * lane], dox[2 * lane + 1] * vx[2 * lane + 1]));
                    const float pxx = exp2f(fmaf(sxx, kLog2e, -lse2_x));
                    if (lane == 0) scal[0] = pxx, scal[1] = pxx * (dpxx - delta_x);
                }
                named_bar_sync(3, kPbEdge * 32);  // the edge warps: row / column vectors and scalars complete
                if (tr) p.trace[cta_id * 32 + 13] = clock64();
                bf16* gx = p.d_qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + h * kHd;
                if (ew == 0) edge_gemv(pcol, sm_do, nv, scal[0], dox, gx + 2 * D, 1.0f, lane);  // dV_x
                if (ew == 1) edge_gemv(dscol, sm_q, nv, scal[1], qx, gx + D, 1.0f, lane);       // dK_x
                if (ew == 2) edge_gemv(dsrow, sm_k, nv, scal[1], kx, gx, 1.0f, lane);           // dQ_x
                __syncwarp();
                // the operand tiles are no longer read from here, and the drain warps may use this item's vectors
                if (lane == 0) mbar_arrive(&bars[kPbEdgeDone]);
                if (tr) p.trace[cta_id * 32 + 12] = clock64();
                // the next item's vectors go into the other slot once the drain warps are done with its last user
                if (it + 1 < n_my) {
                    if (it >= 1) mbar_wait_c(&bars[kPbSlotFree + (s ^ 1)], ((it - 1) >> 1) & 1);
                    prepare_vectors(it + 1);
                }
            }
        } else {
            // ------------------------------------------------------------ drain warps
            const int dw = warp - 9 - kPbEdge;  // 0..3
            const int quarter = warp & 3;        // tensor-memory lane quarter this warp may read (12..15 -> 0..3)
            const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
            uint8_t* stage = sm + kPbOffStage + dw * 4096;
            // delta = rowsum(dO * O) of item `it` into its slot: eight lanes share a 128-byte row, four rows per step,
            // 16 steps per warp in two batches of eight loads in flight
            auto prepare_delta = [&](int it) {
                const int s = it & 1;
                int n, h;
                decode(it, n, h);
                float* delta = vec0 + 1536 * s + 256;
#pragma unroll 1
                for (int base = 0; base < 16; base += 8) {
                    uint4 xo[8], xd[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int row = dw * 64 + (base + u) * 4 + (lane >> 3);
                        const size_t off = (static_cast<size_t>(n) * p.T + min(row, nv)) * D + h * kHd + (lane & 7) * 8;
                        xo[u] = *reinterpret_cast<const uint4*>(p.out + off);
                        xd[u] = *reinterpret_cast<const uint4*>(p.d_out + off);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int row = dw * 64 + (base + u) * 4 + (lane >> 3);
                        float d = chunk_dot(xo[u], xd[u]);
                        d += __shfl_xor_sync(0xffffffffu, d, 1);
                        d += __shfl_xor_sync(0xffffffffu, d, 2);
                        d += __shfl_xor_sync(0xffffffffu, d, 4);
                        if ((lane & 7) == 0) delta[row] = (row < nv) ? d : 0.f;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[kPbVecReady + s]);
            };
            prepare_delta(0);
#pragma unroll 1
            for (int it = 0; it < n_my; ++it) {
                const int s = it & 1;
                const uint32_t ip = it & 1;
                int n, h;
                decode(it, n, h);
                const float* lse2 = vec0 + 1536 * s;
                const float* dscol = lse2 + 768;
                const float* prow = lse2 + 1024;
                const float* dsrow = lse2 + 1280;
                const float* qx = x0 + kPbXFloats * s;
                const float* kx = qx + 64;
                const float* dox = qx + 192;
                const bool tr = (it == kTraceItem) && dw == 0 && lane == 0 && p.trace != nullptr;
                // the next item's delta into the other slot (its last readers, the elementwise warps of the previous
                // item and this warp's own drains, are done)
                if (it + 1 < n_my) prepare_delta(it + 1);
                if (tr) p.trace[cta_id * 32 + 16] = clock64();
                bf16* gd = p.d_qkv + static_cast<size_t>(n) * p.T * 3 * D + h * kHd;
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    mbar_wait_c(&bars[kPbTileDone], (it * 2 + j) & 1);
                    // key tile 0 needs the first half of the edge row; everything after it the whole edge phase
                    mbar_wait_c(&bars[j == 0 ? kPbRowA : kPbEdgeDone], ip);
                    tc_fence_after();
                    if (tr) p.trace[cta_id * 32 + 14 + 4 * j] = clock64();
                    const int row0 = j * 128 + quarter * 32;
                    bwd_drain_tma(trow + kColDV, prow[row0 + lane], dox, stage, &map_dqkv, 2 * D + h * kHd, row0, n, lane);
                    bwd_drain_tma(trow + kColDK, dsrow[row0 + lane], qx, stage, &map_dqkv, D + h * kHd, row0, n, lane);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[kPbAccFreeVK]);
                    if (tr) p.trace[cta_id * 32 + 15 + 4 * j] = clock64();
                }
                mbar_wait_c(&bars[kPbDqDone], ip);
                tc_fence_after();
                bwd_drain_tma(trow + kColDQ, dscol[quarter * 32 + lane], kx, stage, &map_dqkv, h * kHd, quarter * 32, n, lane);
                bwd_drain_tma(trow + kColDQ + 64, dscol[128 + quarter * 32 + lane], kx, stage, &map_dqkv, h * kHd,
                              128 + quarter * 32, n, lane);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[kPbAccFreeQ]), mbar_arrive(&bars[kPbSlotFree + s]);
                if (tr) p.trace[cta_id * 32 + 17] = clock64();
            }
            if (lane == 0) bulk_wait_all();  // this warp's bulk stores have left shared memory and completed
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[cta_id * 32 + 31] = clock64();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// backward, long sequences (T > 257: ViT-L/14 @336)
// ---------------------------------------------------------------------------------------------------------
// The fused backward above keeps one dQ accumulator per query block in tensor memory, which stops at two blocks
// (512 columns).  Here one CTA owns ONE key tile j of a (cutout, head) and walks every query block i:
//   S^T = K_j Q_i^T, dP^T = V_j dO_i^T -> P^T, dS^T (shared memory) -> dV_j += P^T dO_i, dK_j += dS^T Q_i (TMEM,
//   complete when the walk ends) and the PARTIAL dQ_i = dS K_j of this key tile, which is drained from TMEM after
//   every block and reduced across the key-tile CTAs with fp32 vector reductions (red.global.add.v4.f32) into a
//   workspace; a small kernel converts the sums to bf16 afterwards.  Q_i / dO_i stream through a two-stage ring.
// All T tokens are tiled (no edge token: 577 = 4 x 128 + 65); delta = rowsum(dO O) comes from attn_delta_kernel.
// TMEM: S^T 128 | dP^T 128 | dV 64 | dK 64 | dQ partial 64 = 448 of 512 columns, one CTA per SM.
constexpr int kLongThreads = 288;  // warps 0-7 elementwise, warp 8 TMA + MMA
constexpr int kLongMaxT = 1152;
constexpr int kLongOffK = 0;
constexpr int kLongOffV = 1 * kBlkBytes;
constexpr int kLongOffQ = 2 * kBlkBytes;    // two stages
constexpr int kLongOffDO = 4 * kBlkBytes;   // two stages
constexpr int kLongOffPT = 6 * kBlkBytes;   // P^T [128 keys x 128 queries], two 64-column blocks
constexpr int kLongOffDST = 8 * kBlkBytes;  // dS^T, two buffers (block parity): dQ_i = dS K_j reads one while the next block fills the other
constexpr int kLongOffStage = 12 * kBlkBytes;
constexpr int kLongOffVec = 13 * kBlkBytes;  // float lse2[kLongMaxT], delta[kLongMaxT]
constexpr int kLongOffBar = kLongOffVec + 2 * kLongMaxT * 4;
constexpr int kLongSmemBytes = kLongOffBar + 128 + 1024;
constexpr uint32_t kColDQp = 384;
enum { kLbLd = 0, kLbS = 2, kLbA = 3, kLbB = 4, kLbG = 5, kLbD = 6, kLbF = 7 };

struct LongParams {
    int T, heads;
    const float* lse;
    const float* delta;
    float* dq_ws;  // f32 [n*T, D], zeroed by the caller
    bf16* d_qkv;
};

__global__ void __launch_bounds__(kLongThreads, 1)
attn_bwd_long_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                     const LongParams p) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_k = sm + kLongOffK;
    uint8_t* sm_v = sm + kLongOffV;
    uint8_t* sm_q = sm + kLongOffQ;
    uint8_t* sm_do = sm + kLongOffDO;
    uint8_t* sm_pt = sm + kLongOffPT;
    uint8_t* sm_dst = sm + kLongOffDST;
    float* lse2 = reinterpret_cast<float*>(sm + kLongOffVec);
    float* delta = lse2 + kLongMaxT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kLongOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int T = p.T, D = p.heads * kHd;
    const int nq = (T + 127) >> 7;
    const size_t vbase = (static_cast<size_t>(n) * p.heads + h) * T;
    auto width = [&](int i) { return min(128, ((T - 128 * i) + 15) & ~15); };

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(&bars[kLbLd], 1);
            mbar_init(&bars[kLbLd + 1], 1);
            mbar_init(&bars[kLbS], 1);
            mbar_init(&bars[kLbA], 8);
            mbar_init(&bars[kLbB], 8);
            mbar_init(&bars[kLbG], 1);
            mbar_init(&bars[kLbD], 8);
            mbar_init(&bars[kLbF], 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(&bars[kLbLd], 4 * kBlkBytes);
            tma_load_3d(&map_qkv, &bars[kLbLd], sm_k, D + h * kHd, j * 128, n, kEvictNormal);
            tma_load_3d(&map_qkv, &bars[kLbLd], sm_q, h * kHd, 0, n, kEvictNormal);
            tma_load_3d(&map_qkv, &bars[kLbLd], sm_v, 2 * D + h * kHd, j * 128, n, kEvictNormal);
            tma_load_3d(&map_do, &bars[kLbLd], sm_do, h * kHd, 0, n, kEvictNormal);
            if (nq > 1) {
                mbar_arrive_expect_tx(&bars[kLbLd + 1], 2 * kBlkBytes);
                tma_load_3d(&map_qkv, &bars[kLbLd + 1], sm_q + kBlkBytes, h * kHd, 128, n, kEvictNormal);
                tma_load_3d(&map_do, &bars[kLbLd + 1], sm_do + kBlkBytes, h * kHd, 128, n, kEvictNormal);
            }
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    } else {
        // per-query softmax statistics of the whole sequence: lse * log2(e) (+inf past the end: p = 0), delta
        for (int t = threadIdx.x; t < nq * 128; t += 256) {
            lse2[t] = (t < T) ? p.lse[vbase + t] * kLog2e : INFINITY;
            delta[t] = (t < T) ? p.delta[vbase + t] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            // MMA order per block i: [S^T_i, dP^T_i] (first block only; later ones are issued early, see below) ->
            // dV / dK halves as P^T_i / dS^T_i arrive -> S^T_{i+1}, dP^T_{i+1} -> partial dQ_i.  The next block's scores
            // are queued AHEAD of dQ_i so that the elementwise warps start on them while dQ_i (which reads the other
            // dS^T buffer) is still running; the tensor pipe retires in order, so everything that read P^T or the
            // dS^T buffer of block i - 1 is done before S^T_{i+1} signals.
            auto issue_scores = [&](int i) {
                const int st = i & 1;
                mbar_wait(&bars[kLbLd + st], (i >> 1) & 1);
                tc_fence_after();
                mma_tile_x_rows(tmem + kColST, sm_k, sm_q + st * kBlkBytes, width(i));    // S^T  = K_j Q_i^T
                mma_tile_x_rows(tmem + kColDPT, sm_v, sm_do + st * kBlkBytes, width(i));  // dP^T = V_j dO_i^T
                umma_commit(&bars[kLbS]);
            };
            issue_scores(0);
            for (int i = 0; i < nq; ++i) {
                const int st = i & 1;
                const int ksteps = width(i) >> 4, k_lo = min(ksteps, 4);
                const uint8_t* q_i = sm_q + st * kBlkBytes;
                const uint8_t* do_i = sm_do + st * kBlkBytes;
                const uint8_t* dst_i = sm_dst + st * 2 * kBlkBytes;
                if (i >= 1 && i + 1 < nq) {
                    // the other stage held Q / dO of block i - 1: free once dV and dK of that block have retired
                    mbar_wait(&bars[kLbF], (i - 1) & 1);
                    const int s2 = st ^ 1;
                    mbar_arrive_expect_tx(&bars[kLbLd + s2], 2 * kBlkBytes);
                    tma_load_3d(&map_qkv, &bars[kLbLd + s2], sm_q + s2 * kBlkBytes, h * kHd, (i + 1) * 128, n, kEvictNormal);
                    tma_load_3d(&map_do, &bars[kLbLd + s2], sm_do + s2 * kBlkBytes, h * kHd, (i + 1) * 128, n, kEvictNormal);
                }
                mbar_wait(&bars[kLbA], i & 1);
                tc_fence_after();
                mma_blocks_x_cols(tmem + kColDV, sm_pt, do_i, k_lo, i != 0);
                mma_blocks_x_cols(tmem + kColDK, dst_i, q_i, k_lo, i != 0);
                mbar_wait(&bars[kLbB], i & 1);
                tc_fence_after();
                if (ksteps > 4) {
                    mma_blocks_x_cols(tmem + kColDV, sm_pt + kBlkBytes, do_i + 4 * 2048, ksteps - 4, true);
                    mma_blocks_x_cols(tmem + kColDK, dst_i + kBlkBytes, q_i + 4 * 2048, ksteps - 4, true);
                }
                umma_commit(&bars[kLbF]);
                if (i + 1 < nq) issue_scores(i + 1);
                if (i >= 1) {
                    mbar_wait(&bars[kLbD], (i - 1) & 1);  // the previous partial dQ has been drained
                    tc_fence_after();
                }
                mma_rows_t_x_cols(tmem + kColDQp, dst_i, sm_k, 8, false);  // partial dQ_i = dS K_j
                umma_commit(&bars[kLbG]);
            }
        }
    } else {
        const int quarter = warp & 3, phase = warp >> 2;
        const int r = quarter * 32 + lane;  // key row of the tile (S^T, dV, dK) / query row of the block (dQ) == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        uint8_t* stage = sm + kLongOffStage + warp * 2048;
        const bool row_ok = (j * 128 + r) < T;
        // this key tile's share of dQ_i: TMEM -> fp32 vector reductions into the workspace
        auto drain_dq = [&](int i) {
            mbar_wait(&bars[kLbG], i & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld<32>(trow + kColDQp + phase * 32, v);
            tmem_wait_ld();
            const int q = i * 128 + r;
            if (q < T) {
                float* dst = p.dq_ws + (static_cast<size_t>(n) * T + q) * D + h * kHd + phase * 32;
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * g),
                                 "f"(__uint_as_float(v[4 * g])), "f"(__uint_as_float(v[4 * g + 1])),
                                 "f"(__uint_as_float(v[4 * g + 2])), "f"(__uint_as_float(v[4 * g + 3]))
                                 : "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kLbD]);
        };
        for (int i = 0; i < nq; ++i) {
            const int wd = width(i);
            const float* l2 = lse2 + i * 128;
            const float* dl = delta + i * 128;
            uint8_t* dst_i = sm_dst + (i & 1) * 2 * kBlkBytes;
            mbar_wait(&bars[kLbS], i & 1);
            tc_fence_after();
            const int c0 = phase * 32, c1 = 64 + phase * 32;
            uint32_t sa[16], da[16], sb[16], db[16];
            Cols16 o;
            if (c0 < wd) tmem_ld<16>(trow + kColST + c0, sa), tmem_ld<16>(trow + kColDPT + c0, da);
            tmem_wait_ld();
            if (c0 + 16 < wd) tmem_ld<16>(trow + kColST + c0 + 16, sb), tmem_ld<16>(trow + kColDPT + c0 + 16, db);
            if (c0 < wd) {
                o = bwd_cols16(sa, da, l2 + c0, dl + c0, row_ok);
                bwd_store16(o, sm_pt, dst_i, r, c0 >> 3);
            }
            tmem_wait_ld();
            if (c1 < wd) tmem_ld<16>(trow + kColST + c1, sa), tmem_ld<16>(trow + kColDPT + c1, da);
            if (c0 + 16 < wd) {
                o = bwd_cols16(sb, db, l2 + c0 + 16, dl + c0 + 16, row_ok);
                bwd_store16(o, sm_pt, dst_i, r, (c0 + 16) >> 3);
            }
            tmem_wait_ld();
            if (c1 + 16 < wd) tmem_ld<16>(trow + kColST + c1 + 16, sb), tmem_ld<16>(trow + kColDPT + c1 + 16, db);
            if (c1 < wd) o = bwd_cols16(sa, da, l2 + c1, dl + c1, row_ok);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kLbA]);
            if (c1 < wd) bwd_store16(o, sm_pt + kBlkBytes, dst_i + kBlkBytes, r, (c1 - 64) >> 3);
            tmem_wait_ld();
            if (c1 + 16 < wd) {
                o = bwd_cols16(sb, db, l2 + c1 + 16, dl + c1 + 16, row_ok);
                bwd_store16(o, sm_pt + kBlkBytes, dst_i + kBlkBytes, r, (c1 + 16 - 64) >> 3);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kLbB]);
            if (i >= 1) drain_dq(i - 1);  // behind this block's arithmetic: dQ_{i-1} retired while it ran
        }
        drain_dq(nq - 1);
        // dV_j, dK_j are complete (the last wait on kLbG covered every MMA)
        bf16* gd = p.d_qkv + static_cast<size_t>(n) * T * 3 * D + h * kHd + phase * 32;
        const int row0 = j * 128 + quarter * 32;
        bwd_epilogue<false>(trow + kColDV + phase * 32, 0.f, nullptr, stage, lane, gd + 2 * D, static_cast<size_t>(3) * D,
                            row0, T);
        bwd_epilogue<false>(trow + kColDK + phase * 32, 0.f, nullptr, stage, lane, gd + D, static_cast<size_t>(3) * D, row0,
                            T);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

// dQ sums (f32 [rows, D]) -> the q third of d_qkv (bf16 [rows, 3D]); 8 elements per thread
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ dq, bf16* __restrict__ d_qkv,
                                                              size_t rows, int D) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int per_row = D >> 3;
    if (idx >= rows * per_row) return;
    const size_t row = idx / per_row;
    const int c = static_cast<int>(idx - row * per_row) * 8;
    const float4 a = *reinterpret_cast<const float4*>(dq + row * D + c);
    const float4 b = *reinterpret_cast<const float4*>(dq + row * D + c + 4);
    *reinterpret_cast<uint4*>(d_qkv + row * 3 * D + c) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

// ---------------------------------------------------------------------------------------------------------
// short sequences (T <= 64: ViT-B/32 has 49 + 1 tokens): two heads per 128-row tile
// ---------------------------------------------------------------------------------------------------------
// A 50-token head fills 39 % of a 128-row MMA tile, so one CTA packs heads (2g, 2g+1) of a cutout: rows / keys
// [0, 64) belong to the first head, [64, 128) to the second (tokens >= T are zero rows from the TMA bounds check).
// S = Q K^T is then a 128 x 128 tile of which only the two diagonal 64 x 64 blocks mean anything; P is written
// block-diagonal (zeros elsewhere), so P V, P^T dO, dS^T Q and dS K never mix the heads.  The tensor work is
// trivial (12 / 32 MMAs per pair); these kernels exist to replace ~40 small dependent mma.sync steps per head by
// one TMA round trip and a handful of tensor-memory operations.
constexpr int kPackFwdThreads = 160;  // warps 0-3 softmax (row == TMEM lane), warp 4 TMA + MMA
constexpr int kPackOffK = 0, kPackOffV = kBlkBytes, kPackOffQ = 2 * kBlkBytes;
constexpr int kPackFwdOffBar = 3 * kBlkBytes;
constexpr int kPackFwdSmem = kPackFwdOffBar + 64 + 1024;
constexpr uint32_t kPackColO = 128;

struct PackParams {
    int T, heads;
    const bf16* qkv;
    bf16* out;
    float* lse;
    // backward only
    const bf16* d_out;
    const float* delta;
    bf16* d_qkv;
};

// both heads' [64 x 64] boxes of one operand (column offset col0 of head h0) into a [128 x 64] tile
__device__ __forceinline__ void pack_load(const CUtensorMap* map, uint64_t* bar, uint8_t* tile, int col0, int n) {
    tma_load_3d(map, bar, tile, col0, 0, n, kEvictFirst);
    tma_load_3d(map, bar, tile + 64 * 128, col0 + kHd, 0, n, kEvictFirst);
}

__global__ void __launch_bounds__(kPackFwdThreads, 2)
attn_fwd_pack_kernel(const __grid_constant__ CUtensorMap map_qkv, const PackParams p) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_k = sm + kPackOffK;
    uint8_t* sm_v = sm + kPackOffV;
    uint8_t* sm_q = sm + kPackOffQ;
    // mbarriers: 0 operands landed, 1 S ready, 2 P stored, 3 O ready
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kPackFwdOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h0 = 2 * blockIdx.x, n = blockIdx.y;
    const int T = p.T, D = p.heads * kHd;

    if (warp == 4) {
        if (lane == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            mbar_init(&bars[2], 4);
            mbar_init(&bars[3], 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(&bars[0], 3 * kBlkBytes);
            pack_load(&map_qkv, &bars[0], sm_q, h0 * kHd, n);
            pack_load(&map_qkv, &bars[0], sm_k, D + h0 * kHd, n);
            pack_load(&map_qkv, &bars[0], sm_v, 2 * D + h0 * kHd, n);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            mbar_wait(&bars[0], 0);
            tc_fence_after();
            mma_tile_x_rows(tmem, sm_q, sm_k, 128);  // S = Q K^T, both heads at once
            umma_commit(&bars[1]);
            mbar_wait(&bars[2], 0);
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
            for (int ks = 0; ks < 8; ++ks)
                umma_f16_ts(tmem + kPackColO, tmem + ks * 8, umma_smem_desc_sw128(smem_u32(sm_v + ks * 2048)), idesc, ks != 0);
            umma_commit(&bars[3]);
        }
    } else {
        const int r = warp * 32 + lane;  // tile row == TMEM lane
        const int blk = r >> 6, t = r & 63;
        const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        uint32_t v[64];
        tmem_ld_32x32(trow + 64 * blk, reinterpret_cast<uint32_t(&)[32]>(v[0]));
        tmem_ld_32x32(trow + 64 * blk + 32, reinterpret_cast<uint32_t(&)[32]>(v[32]));
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; ++c) mx = fmaxf(mx, (c < T) ? __uint_as_float(v[c]) : -INFINITY);
        const float mb = mx * kLog2e;
        float sum = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float e0 = (2 * c < T) ? exp2f(fmaf(__uint_as_float(v[2 * c]), kLog2e, -mb)) : 0.f;
            const float e1 = (2 * c + 1 < T) ? exp2f(fmaf(__uint_as_float(v[2 * c + 1]), kLog2e, -mb)) : 0.f;
            sum += e0 + e1;
            pk[c] = pack_bf16(e0, e1);
        }
        // P row: this head's 64 keys as 32 packed columns, zeros for the other head's keys
        uint32_t zero[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) zero[c] = 0u;
        tmem_st<16>(trow + 32 * blk, reinterpret_cast<uint32_t(&)[16]>(pk[0]));
        tmem_st<16>(trow + 32 * blk + 16, reinterpret_cast<uint32_t(&)[16]>(pk[16]));
        tmem_st<16>(trow + 32 * (1 - blk), zero);
        tmem_st<16>(trow + 32 * (1 - blk) + 16, zero);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2]);
        const int h = h0 + blk;
        if (t < T) p.lse[(static_cast<size_t>(n) * p.heads + h) * T + t] = mx + logf(sum);
        const float inv = 1.0f / sum;
        mbar_wait(&bars[3], 0);
        tc_fence_after();
        tmem_ld_32x32(trow + kPackColO, reinterpret_cast<uint32_t(&)[32]>(v[0]));
        tmem_ld_32x32(trow + kPackColO + 32, reinterpret_cast<uint32_t(&)[32]>(v[32]));
        tmem_wait_ld();
        if (t < T) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(n) * T + t) * D + h * kHd);
#pragma unroll
            for (int g = 0; g < 8; ++g)
                dst[g] = make_uint4(pack_bf16(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv),
                                    pack_bf16(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv),
                                    pack_bf16(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv),
                                    pack_bf16(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 256);
}

constexpr int kPackBwdThreads = 288;  // warps 0-7 elementwise (lane quarter = warp & 3, column half = warp >> 2), 8 TMA + MMA
constexpr int kPackOffDO = 3 * kBlkBytes;
constexpr int kPackOffDST = 4 * kBlkBytes;     // dS^T [128 keys x 128 queries], two 64-column blocks
constexpr int kPackBwdOffVec = 6 * kBlkBytes;  // float lse2[128], delta[128]
constexpr int kPackBwdOffBar = kPackBwdOffVec + 1024;
constexpr int kPackBwdSmem = kPackBwdOffBar + 64 + 1024;
// 256 TMEM columns so that two CTAs share an SM: S^T [0,128) and dP^T [128,256) are dead once the elementwise warps
// have read them, so P^T (packed bf16, the A operand of dV) goes over S^T [0,64), dQ over S^T [64,128), dV and dK
// over dP^T.
constexpr uint32_t kPkST = 0, kPkDPT = 128, kPkPT = 0, kPkDQ = 64, kPkDV = 128, kPkDK = 192;

__global__ void __launch_bounds__(kPackBwdThreads, 2)
attn_bwd_pack_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                     const PackParams p) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_k = sm + kPackOffK;
    uint8_t* sm_v = sm + kPackOffV;
    uint8_t* sm_q = sm + kPackOffQ;
    uint8_t* sm_do = sm + kPackOffDO;
    uint8_t* sm_dst = sm + kPackOffDST;
    float* lse2 = reinterpret_cast<float*>(sm + kPackBwdOffVec);
    float* delta = lse2 + 128;
    // mbarriers: 0 operands landed, 1 S^T / dP^T ready, 2 P^T / dS^T stored, 3 gradients ready
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kPackBwdOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h0 = 2 * blockIdx.x, n = blockIdx.y;
    const int T = p.T, D = p.heads * kHd;

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            mbar_init(&bars[2], 8);
            mbar_init(&bars[3], 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(&bars[0], 4 * kBlkBytes);
            pack_load(&map_qkv, &bars[0], sm_k, D + h0 * kHd, n);
            pack_load(&map_qkv, &bars[0], sm_q, h0 * kHd, n);
            pack_load(&map_qkv, &bars[0], sm_v, 2 * D + h0 * kHd, n);
            pack_load(&map_do, &bars[0], sm_do, h0 * kHd, n);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    } else if (threadIdx.x < 128) {
        const int blk = threadIdx.x >> 6, t = threadIdx.x & 63;
        const size_t idx = (static_cast<size_t>(n) * p.heads + h0 + blk) * T + min(t, T - 1);
        lse2[threadIdx.x] = (t < T) ? p.lse[idx] * kLog2e : INFINITY;  // +inf: p = 0 for queries past the end
        delta[threadIdx.x] = (t < T) ? p.delta[idx] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            mbar_wait(&bars[0], 0);
            tc_fence_after();
            mma_tile_x_rows(tmem + kPkST, sm_k, sm_q, 128);    // S^T  = K Q^T
            mma_tile_x_rows(tmem + kPkDPT, sm_v, sm_do, 128);  // dP^T = V dO^T
            umma_commit(&bars[1]);
            mbar_wait(&bars[2], 0);
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
            for (int ks = 0; ks < 8; ++ks)  // dV = P^T dO, A = P^T from tensor memory
                umma_f16_ts(tmem + kPkDV, tmem + kPkPT + ks * 8, umma_smem_desc_sw128(smem_u32(sm_do + ks * 2048)), idesc,
                            ks != 0);
            mma_blocks_x_cols(tmem + kPkDK, sm_dst, sm_q, 8, false);  // dK = dS^T Q
            mma_rows_t_x_cols(tmem + kPkDQ, sm_dst, sm_k, 8, false);  // dQ = dS K
            umma_commit(&bars[3]);
        }
    } else {
        const int quarter = warp & 3, half = warp >> 2;
        const int r = quarter * 32 + lane;  // key row (S^T, dV, dK) / query row (dQ) == TMEM lane
        const int blk = r >> 6, t = r & 63;
        const bool row_ok = t < T;
        const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        // this head's 64 query columns: this warp takes 32 of them, 16 at a time
        const int c0 = 64 * blk + 32 * half;
        uint32_t sa[16], da[16], sb[16], db[16];
        tmem_ld<16>(trow + kPkST + c0, sa), tmem_ld<16>(trow + kPkDPT + c0, da);
        tmem_ld<16>(trow + kPkST + c0 + 16, sb), tmem_ld<16>(trow + kPkDPT + c0 + 16, db);
        // the other head's columns of this row are zero (block-diagonal P^T, dS^T): this warp clears half of them
        {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            uint8_t* zd = sm_dst + (1 - blk) * kBlkBytes;
#pragma unroll
            for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(zd + row_chunk(r, 4 * half + g)) = z;
        }
        tmem_wait_ld();
        // P^T overwrites S^T columns that the OTHER column-half warp of this lane quarter may still be reading
        named_bar_sync(1, 256);
        const Cols16 oa = bwd_cols16(sa, da, lse2 + c0, delta + c0, row_ok);
        const Cols16 ob = bwd_cols16(sb, db, lse2 + c0 + 16, delta + c0 + 16, row_ok);
        bwd_store_ds(oa, sm_dst + blk * kBlkBytes, r, (c0 & 63) >> 3);
        bwd_store_ds(ob, sm_dst + blk * kBlkBytes, r, ((c0 + 16) & 63) >> 3);
        {
            // packed P^T row: queries of head 0 in columns [0, 32), of head 1 in [32, 64)
            const uint32_t pa[8] = {oa.p[0].x, oa.p[0].y, oa.p[0].z, oa.p[0].w, oa.p[1].x, oa.p[1].y, oa.p[1].z, oa.p[1].w};
            const uint32_t pb[8] = {ob.p[0].x, ob.p[0].y, ob.p[0].z, ob.p[0].w, ob.p[1].x, ob.p[1].y, ob.p[1].z, ob.p[1].w};
            uint32_t zero[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) zero[c] = 0u;
            tmem_st<8>(trow + kPkPT + 32 * blk + 16 * half, pa);
            tmem_st<8>(trow + kPkPT + 32 * blk + 16 * half + 8, pb);
            tmem_st<16>(trow + kPkPT + 32 * (1 - blk) + 16 * half, zero);
        }
        tmem_wait_st();
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2]);
        mbar_wait(&bars[3], 0);
        tc_fence_after();
        bf16* grow = p.d_qkv + (static_cast<size_t>(n) * T + min(t, T - 1)) * 3 * D + (h0 + blk) * kHd + 32 * half;
        const uint32_t cols[3] = {kPkDQ, kPkDK, kPkDV};
#pragma unroll
        for (int which = 0; which < 3; ++which) {
            uint32_t v[32];
            tmem_ld<32>(trow + cols[which] + 32 * half, v);  // warp-collective: every lane takes part, valid rows store
            tmem_wait_ld();
            if (row_ok) {
                uint4* dst = reinterpret_cast<uint4*>(grow + which * D);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    dst[g] = make_uint4(pack_bf16(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                                        pack_bf16(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                                        pack_bf16(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                                        pack_bf16(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 256);
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

// [n][T][cols] bf16 view of a row-major [n*T, cols] matrix; box = {64 columns, box_rows (128 or 64) rows, 1}.
// Rows >= T of a box are zero-filled, so whole boxes are always loaded.
// `rows_dim` (default T): the extent of the row dimension when it is shorter than the stride between cutouts -- a store
// map that must not touch the last token's row.
int make_map3(CUtensorMap* out, const void* ptr, int n, int T, int cols, int box_rows = 128, int rows_dim = 0) {
    if (rows_dim <= 0) rows_dim = T;
    struct Key {
        const void* p;
        int n, T, cols, box, rows;
        bool operator==(const Key& o) const {
            return p == o.p && n == o.n && T == o.T && cols == o.cols && box == o.box && rows == o.rows;
        }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            size_t h = reinterpret_cast<size_t>(k.p);
            for (int v : {k.n, k.T, k.cols, k.box, k.rows}) h = h * 1000003u ^ static_cast<size_t>(v);
            return h;
        }
    };
    static std::mutex mu;
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    const Key key{ptr, n, T, cols, box_rows, rows_dim};
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
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows_dim), static_cast<cuuint64_t>(n)};
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

// tensor-core path: at least 65 tokens beside the edge token so that a 128-row tile is not mostly padding.
// T - 1 <= 256: everything in tensor memory at once (forward) / one fused backward; beyond that the streaming forward.
bool use_tc(int T) { return T >= 66 && T <= 257; }
bool use_flash_fwd(int T) { return T > 257; }
bool g_pack_enabled = []() {  // PCG_ATTN_PACK=0 sends T <= 64 back to the mma.sync kernels
    const char* e = getenv("PCG_ATTN_PACK");
    return !(e != nullptr && e[0] == '0');
}();
bool use_pack(int T, int heads) { return g_pack_enabled && T <= 64 && heads % 2 == 0; }  // two heads per 128-row tile
bool use_long_bwd(int T) { return T > 257 && T <= kLongMaxT; }
size_t bwd_delta_floats(int n, int T, int heads) { return align_up(static_cast<size_t>(n) * heads * T, 64); }

long long* g_trace = nullptr;
int g_bwd_stagger = []() {
    const char* e = getenv("PCG_ATTN_BWD_STAGGER");
    return e != nullptr ? atoi(e) : 2500;  // measured: 279 us (0) -> 274 us (2500) per ViT-L/14 layer, 6000 is worse
}();
int g_fwd_stagger = []() {
    const char* e = getenv("PCG_ATTN_STAGGER");
    return e != nullptr ? atoi(e) : 5000;
}();

// PCG_ATTN_PERSIST=0 sends 66 <= T <= 257 back to the one-tile-per-CTA kernels (kept for A/B measurements and tests)
bool g_fwd_persist = []() {
    const char* e = getenv("PCG_ATTN_PERSIST");
    return !(e != nullptr && e[0] == '0');
}();

// PCG_ATTN_SPLIT=0 keeps the four-softmax-warp persistent forward for T - 1 > 128 too (A/B)
bool g_fwd_split = []() {
    const char* e = getenv("PCG_ATTN_SPLIT");
    return !(e != nullptr && e[0] == '0');
}();

bool g_bwd_persist = []() {
    const char* e = getenv("PCG_ATTN_PERSIST");
    return !(e != nullptr && e[0] == '0');
}();

bool g_force_legacy = []() {
    const char* e = getenv("PCG_ATTN_LEGACY");
    return e != nullptr && e[0] == '1';
}();

}  // namespace
}  // namespace pcg

using namespace pcg;

extern "C" int pcg_attn_set_legacy(int on) {  // test hook: force the mma.sync kernels for every row
    g_force_legacy = on != 0;
    return 0;
}

// test / benchmark hook: 0 one-tile-per-CTA tcgen05 kernels, 1 persistent forward + backward (the default), 2 persistent
// backward only, 3 persistent forward only
extern "C" int pcg_attn_set_persist(int mode) {
    g_fwd_persist = mode == 1 || mode == 3;
    g_bwd_persist = mode == 1 || mode == 2;
    return 0;
}

extern "C" int pcg_attn_set_split(int on) {  // test / benchmark hook: 0 = four softmax warps per forward CTA, 1 = eight
    g_fwd_split = on != 0;
    return 0;
}

extern "C" int pcg_attn_set_trace(void* device_buf) {  // profiling hook: [ctas][32] int64, see tools/attn_trace.py
    g_trace = static_cast<long long*>(device_buf);
    return 0;
}

extern "C" int pcg_attn_fwd(const void* qkv, void* out, float* lse, int n, int T, int heads, void* stream) {
    PCG_CHECK_ARG(qkv && out && lse, "pcg_attn_fwd: null pointer");
    PCG_CHECK_ARG(n > 0 && T > 0 && heads > 0, "pcg_attn_fwd: bad shape n=%d T=%d heads=%d", n, T, heads);
    PCG_CHECK_ARG(n <= 65535 && heads <= 65535, "pcg_attn_fwd: n and heads must be <= 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_FWD, 4.0 * T * T * kHd * heads * n, s);
    const int D = heads * kHd;
    if (!g_force_legacy && use_pack(T, heads)) {
        CUtensorMap pmap;
        if (int rc = make_map3(&pmap, qkv, n, T, 3 * D, 64)) return rc;
        static PerDeviceOnce pack_configured;
        PCG_ONCE_PER_DEVICE(pack_configured, PCG_CUDA(cudaFuncSetAttribute(attn_fwd_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPackFwdSmem)));
        PackParams pp{T, heads, static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, nullptr, nullptr, nullptr};
        attn_fwd_pack_kernel<<<dim3(heads / 2, n), kPackFwdThreads, kPackFwdSmem, s>>>(pmap, pp);
        PCG_LAUNCH_CHECK("attn_fwd_pack_kernel");
        return 0;
    }
    if (g_force_legacy || !(use_tc(T) || use_flash_fwd(T))) return attn_fwd_legacy(qkv, out, lse, n, T, heads, 0, s);
    CUtensorMap map;
    if (int rc = make_map3(&map, qkv, n, T, 3 * D)) return rc;
    if (use_flash_fwd(T)) {
        static PerDeviceOnce flash_configured;
        PCG_ONCE_PER_DEVICE(flash_configured, PCG_CUDA(cudaFuncSetAttribute(attn_fwd_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlashSmemBytes)));
        FwdParams pf{T, heads, T, (T + 15) & ~15, static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse,
                     nullptr, 2 * sm_count(), g_fwd_stagger};
        attn_fwd_flash_kernel<<<dim3((T + 127) / 128, heads, n), kFwdThreads, kFlashSmemBytes, s>>>(map, pf);
        PCG_LAUNCH_CHECK("attn_fwd_flash_kernel");
        return 0;
    }
    const int nv = T - 1;
    // two key tiles: every score row split over two warps.  (The size limits keep the kernel's multiply-high quotients exact.)
    if (g_fwd_persist && g_fwd_split && nv > 128 && heads >= 2 && heads <= 1024 && static_cast<long long>(n) * heads * 2 < (1ll << 22)) {
        static PerDeviceOnce split_configured;
        PCG_ONCE_PER_DEVICE(split_configured, PCG_CUDA(cudaFuncSetAttribute(attn_fwd_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd3SmemBytes)));
        const int items = n * heads * 2;
        const int grid = std::min(items, 2 * sm_count());
        Fwd3Params p3;
        static_cast<Fwd2Params&>(p3) = Fwd2Params{T, heads, nv, (nv + 15) & ~15, 2, items, static_cast<const bf16*>(qkv),
                                                  static_cast<bf16*>(out), lse, g_trace, 0};
        p3.heads_magic = static_cast<uint32_t>(((1ull << 32) + heads - 1) / heads);
        p3.grid_magic = static_cast<uint32_t>(((1ull << 32) + grid - 1) / grid);
        // store map of the output: rows [0, nv) of every cutout (the edge token's row is the edge warp's), 32-row boxes
        CUtensorMap map_out;
        if (int rc = make_map3(&map_out, out, n, T, D, 32, nv)) return rc;
        attn_fwd_split_kernel<<<grid, kFwd3Threads, kFwd3SmemBytes, s>>>(map, map_out, p3);
        PCG_LAUNCH_CHECK("attn_fwd_split_kernel");
        return 0;
    }
    if (g_fwd_persist) {
        static PerDeviceOnce persist_configured;
        PCG_ONCE_PER_DEVICE(persist_configured, PCG_CUDA(cudaFuncSetAttribute(attn_fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd2SmemBytes)));
        const int tiles = (nv + 127) / 128;
        const long long items = static_cast<long long>(n) * heads * tiles;
        PCG_CHECK_ARG(items < (1ll << 30), "pcg_attn_fwd: too many (cutout, head, tile) items");
        Fwd2Params p2{T, heads, nv, (nv + 15) & ~15, tiles, static_cast<int>(items), static_cast<const bf16*>(qkv),
                      static_cast<bf16*>(out), lse, g_trace, g_fwd_stagger};
        const int grid = static_cast<int>(std::min<long long>(items, 2ll * sm_count()));
        attn_fwd_persist_kernel<<<grid, kFwdThreads, kFwd2SmemBytes, s>>>(map, p2);
        PCG_LAUNCH_CHECK("attn_fwd_persist_kernel");
        return 0;
    }
    static PerDeviceOnce configured;
    PCG_ONCE_PER_DEVICE(configured, PCG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes)));
    FwdParams p{T, heads, nv, (nv + 15) & ~15, static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse,
                g_trace, 2 * sm_count(), g_fwd_stagger};
    attn_fwd_tc_kernel<<<dim3((nv + 127) / 128, heads, n), kFwdThreads, kFwdSmemBytes, s>>>(map, p);
    PCG_LAUNCH_CHECK("attn_fwd_tc_kernel");
    return 0;
}

extern "C" size_t pcg_attn_bwd_workspace_bytes(int n, int T, int heads) {
    if (n <= 0 || T <= 0 || heads <= 0) return 0;
    size_t floats = bwd_delta_floats(n, T, heads);
    if (use_long_bwd(T)) floats += static_cast<size_t>(n) * T * heads * kHd;  // fp32 dQ sums of the key-tile CTAs
    return floats * sizeof(float);
}

extern "C" int pcg_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                            void* d_qkv, int n, int T, int heads, void* stream) {
    PCG_CHECK_ARG(qkv && out && d_out && lse && delta_ws && d_qkv, "pcg_attn_bwd: null pointer");
    PCG_CHECK_ARG(n > 0 && T > 0 && heads > 0, "pcg_attn_bwd: bad shape n=%d T=%d heads=%d", n, T, heads);
    PCG_CHECK_ARG(n <= 65535 && heads <= 65535, "pcg_attn_bwd: n and heads must be <= 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_BWD, 8.0 * T * T * kHd * heads * n, s);
    const int D = heads * kHd;
    if (!g_force_legacy && use_pack(T, heads)) {
        CUtensorMap pmap, pmap_do;
        if (int rc = make_map3(&pmap, qkv, n, T, 3 * D, 64)) return rc;
        if (int rc = make_map3(&pmap_do, d_out, n, T, D, 64)) return rc;
        static PerDeviceOnce pack_configured;
        PCG_ONCE_PER_DEVICE(pack_configured, PCG_CUDA(cudaFuncSetAttribute(attn_bwd_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPackBwdSmem)));
        if (int rc = attn_delta(out, d_out, delta_ws, n, T, heads, s)) return rc;
        PackParams pp{T, heads, static_cast<const bf16*>(qkv), nullptr, const_cast<float*>(lse),
                      static_cast<const bf16*>(d_out), delta_ws, static_cast<bf16*>(d_qkv)};
        attn_bwd_pack_kernel<<<dim3(heads / 2, n), kPackBwdThreads, kPackBwdSmem, s>>>(pmap, pmap_do, pp);
        PCG_LAUNCH_CHECK("attn_bwd_pack_kernel");
        return 0;
    }
    if (!g_force_legacy && use_long_bwd(T)) {
        CUtensorMap lmap, lmap_do;
        if (int rc = make_map3(&lmap, qkv, n, T, 3 * D)) return rc;
        if (int rc = make_map3(&lmap_do, d_out, n, T, D)) return rc;
        static PerDeviceOnce long_configured;
        PCG_ONCE_PER_DEVICE(long_configured, PCG_CUDA(cudaFuncSetAttribute(attn_bwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLongSmemBytes)));
        if (int rc = attn_delta(out, d_out, delta_ws, n, T, heads, s)) return rc;
        float* dq_ws = delta_ws + bwd_delta_floats(n, T, heads);
        const size_t rows = static_cast<size_t>(n) * T;
        PCG_CUDA(cudaMemsetAsync(dq_ws, 0, rows * D * sizeof(float), s));
        LongParams lp{T, heads, lse, delta_ws, dq_ws, static_cast<bf16*>(d_qkv)};
        attn_bwd_long_kernel<<<dim3((T + 127) / 128, heads, n), kLongThreads, kLongSmemBytes, s>>>(lmap, lmap_do, lp);
        PCG_LAUNCH_CHECK("attn_bwd_long_kernel");
        const size_t items = rows * (D / 8);
        attn_dq_convert_kernel<<<static_cast<unsigned>((items + 255) / 256), 256, 0, s>>>(dq_ws, static_cast<bf16*>(d_qkv),
                                                                                         rows, D);
        PCG_LAUNCH_CHECK("attn_dq_convert_kernel");
        return 0;
    }
    if (g_force_legacy || !use_tc(T)) {
        if (int rc = attn_delta(out, d_out, delta_ws, n, T, heads, s)) return rc;
        return attn_bwd_legacy(qkv, d_out, lse, delta_ws, d_qkv, n, T, heads, 0, s);
    }
    CUtensorMap map, map_do;
    if (int rc = make_map3(&map, qkv, n, T, 3 * D)) return rc;
    if (int rc = make_map3(&map_do, d_out, n, T, D)) return rc;
    const int nv = T - 1;
    if (g_bwd_persist && nv > 128) {  // two key tiles: the persistent kernel (a single tile keeps the one-shot kernel)
        static PerDeviceOnce persist_configured;
        PCG_ONCE_PER_DEVICE(persist_configured, PCG_CUDA(cudaFuncSetAttribute(attn_bwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPbSmemBytes)));
        const long long items = static_cast<long long>(n) * heads;
        PCG_CHECK_ARG(items < (1ll << 30), "pcg_attn_bwd: too many (cutout, head) items");
        PbParams pb{T, heads, nv, static_cast<int>(items), static_cast<const bf16*>(qkv), static_cast<const bf16*>(out),
                    static_cast<const bf16*>(d_out), lse, static_cast<bf16*>(d_qkv), g_trace};
        // store map of d_qkv: rows [0, nv) of every cutout (the edge token's row is written by the edge warps), 32-row boxes
        CUtensorMap map_dqkv;
        if (int rc = make_map3(&map_dqkv, d_qkv, n, T, 3 * D, 32, nv)) return rc;
        const int grid = static_cast<int>(std::min<long long>(items, sm_count()));
        attn_bwd_persist_kernel<<<grid, kPbThreads, kPbSmemBytes, s>>>(map, map_do, map_dqkv, pb);
        PCG_LAUNCH_CHECK("attn_bwd_persist_kernel");
        return 0;
    }
    static PerDeviceOnce configured;
    PCG_ONCE_PER_DEVICE(configured, PCG_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes)));
    BwdParams p{T,   heads, nv, (nv + 127) / 128, static_cast<const bf16*>(qkv), static_cast<const bf16*>(out),
                static_cast<const bf16*>(d_out), lse, static_cast<bf16*>(d_qkv), g_trace, sm_count(), g_bwd_stagger};  // delta: in-kernel
    attn_bwd_tc_kernel<<<dim3(heads, n), kBwdThreads, kBwdSmemBytes, s>>>(map, map_do, p);
    PCG_LAUNCH_CHECK("attn_bwd_tc_kernel");
    return 0;
}
