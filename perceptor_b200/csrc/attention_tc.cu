// tcgen05 short-sequence attention for 66 <= T <= 257 (ViT-B/16: 197, ViT-L/14: 257), head dim 64: forward and a
// single fused backward (dQ, dK, dV in one pass, S and dP computed once).
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
__device__ __forceinline__ float row_dot(const uint8_t* mat, int row, const float* vec) {
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
__device__ __forceinline__ void edge_gemv(const float* coef, const uint8_t* mat, int rows, float corner,
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
#define PCG_TRACE(slot)                                                       \
    do {                                                                      \
        if (p.trace != nullptr && lane == 0) p.trace[cta_id * 32 + (slot)] = clock64(); \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 192;  // warps 0-3 softmax (TMEM lane quarters), 4 TMA + MMA, 5 edge query row
constexpr int kFwdOffV = 2 * kBlkBytes;
constexpr int kFwdOffQ = 4 * kBlkBytes;
constexpr int kFwdOffX = 5 * kBlkBytes;               // float k_x[64], v_x[64], q_x[64], p_x[256]
constexpr int kFwdOffBar = kFwdOffX + 3 * 256 + 1024;  // 5 mbarriers + tmem slot
constexpr int kFwdSmemBytes = kFwdOffBar + 64 + 1024;
constexpr uint32_t kFwdColO = 128;

struct FwdParams {
    int T, heads;
    int nv;  // T - 1: tokens on the tensor cores (as queries and as keys); token nv is the edge token
    int nk;  // nv rounded up to 16: MMA N of S / K extent of P V
    const bf16* qkv;
    bf16* out;
    float* lse;
    long long* trace;  // optional [ctas][32] clock64 stamps (tools/attn_trace.py), nullptr in production
};

template <int W>
__device__ __forceinline__ float fwd_chunk_max(uint32_t taddr, int c, int nv, float mx) {
    uint32_t v[W];
    tmem_ld<W>(taddr + c, v);
    tmem_wait_ld();
    if (c + W <= nv) {
#pragma unroll
        for (int j = 0; j < W; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    } else {
#pragma unroll
        for (int j = 0; j < W; ++j) mx = fmaxf(mx, (c + j < nv) ? __uint_as_float(v[j]) : -INFINITY);
    }
    return mx;
}
// P = exp2(S log2e - mb) for W columns, written back over the (already consumed) S columns as packed bf16
template <int W>
__device__ __forceinline__ float fwd_chunk_exp(uint32_t taddr, int c, int nv, float mb, float sum) {
    uint32_t v[W];
    tmem_ld<W>(taddr + c, v);
    tmem_wait_ld();
    uint32_t pk[W / 2];
    const bool full = c + W <= nv;
#pragma unroll
    for (int j = 0; j < W / 2; ++j) {
        float e0 = exp2f(fmaf(__uint_as_float(v[2 * j]), kLog2e, -mb));
        float e1 = exp2f(fmaf(__uint_as_float(v[2 * j + 1]), kLog2e, -mb));
        if (!full) {
            e0 = (c + 2 * j < nv) ? e0 : 0.f;
            e1 = (c + 2 * j + 1 < nv) ? e1 : 0.f;
        }
        sum += e0 + e1;
        pk[j] = pack_bf16(e0, e1);
    }
    tmem_st<W / 2>(taddr + (c >> 1), pk);
    return sum;
}

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const FwdParams p) {
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
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kFwdOffBar);  // 0 K+Q, 1 V, 2 S, 3 P, 4 O
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int D = p.heads * kHd, nv = p.nv, nk = p.nk;
    const int q0 = tile * 128;
    const bool has_edge_row = (q0 + 128 >= nv);  // the CTA of the last tile also computes query row x
    const size_t cta_id = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (warp == 0) PCG_TRACE(0);
    const bf16* xrow_g = p.qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + h * kHd;

    if (warp == 4) {
        if (lane == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            mbar_init(&bars[2], 1);
            mbar_init(&bars[3], 4);
            mbar_init(&bars[4], 1);
            fence_barrier_init();
            tma_prefetch_desc(&map_qkv);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    } else if (warp == 5) {
        load_row_f32(qx, xrow_g, lane);
        load_row_f32(kx, xrow_g + D, lane);
        load_row_f32(vx, xrow_g + 2 * D, lane);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) PCG_TRACE(1);

    if (warp == 4) {
        if (lane == 0) {
            const int nblk = (nk + 127) >> 7;
            mbar_arrive_expect_tx(&bars[0], (nblk + 1) * kBlkBytes);
            tma_load_3d(&map_qkv, &bars[0], sm_q, h * kHd, q0, n, kEvictFirst);
            for (int i = 0; i < nblk; ++i)
                tma_load_3d(&map_qkv, &bars[0], sm_k + i * kBlkBytes, D + h * kHd, i * 128, n, kEvictNormal);
            mbar_arrive_expect_tx(&bars[1], nblk * kBlkBytes);
            for (int i = 0; i < nblk; ++i)
                tma_load_3d(&map_qkv, &bars[1], sm_v + i * kBlkBytes, 2 * D + h * kHd, i * 128, n, kEvictNormal);
            mbar_wait(&bars[0], 0);
            PCG_TRACE(2);
            tc_fence_after();
            mma_tile_x_rows(tmem, sm_q, sm_k, nk);  // S = Q K^T
            umma_commit(&bars[2]);
            mbar_wait(&bars[1], 0);  // V landed
            PCG_TRACE(3);
            mbar_wait(&bars[3], 0);  // P is in TMEM
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
            const int ksteps = nk >> 4;
            for (int ks = 0; ks < ksteps; ++ks)  // O = P V, A = P from TMEM (8 columns of bf16 pairs per step)
                umma_f16_ts(tmem + kFwdColO, tmem + ks * 8, umma_smem_desc_sw128(smem_u32(sm_v + ks * 2048)), idesc,
                            ks != 0);
            umma_commit(&bars[4]);
        }
    } else if (warp < 4) {
        const int r = warp * 32 + lane;  // query row in the tile == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        mbar_wait(&bars[0], 0);
        const float sx = row_dot(sm_q, r, kx);  // score against the edge key, while the MMA runs
        mbar_wait(&bars[2], 0);
        tc_fence_after();
        if (warp == 0) PCG_TRACE(4);
        float mx = sx;
        int c = 0;
        for (; c + 32 <= nk; c += 32) mx = fwd_chunk_max<32>(trow, c, nv, mx);
        if (c < nk) mx = fwd_chunk_max<16>(trow, c, nv, mx);
        const float mb = mx * kLog2e;
        float sum = 0.f;
        if (warp == 0) PCG_TRACE(5);
        for (c = 0; c + 32 <= nk; c += 32) sum = fwd_chunk_exp<32>(trow, c, nv, mb, sum);
        if (c < nk) sum = fwd_chunk_exp<16>(trow, c, nv, mb, sum);
        const float px = exp2f(fmaf(sx, kLog2e, -mb));
        sum += px;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[3]);
        if (warp == 0) PCG_TRACE(6);
        if (q0 + r < nv) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + q0 + r] = mx + logf(sum);
        const float inv = 1.0f / sum;
        mbar_wait(&bars[4], 0);
        tc_fence_after();
        if (warp == 0) PCG_TRACE(7);
        // O row: (P V + p_x v_x) / sum -> bf16 through the (now free) Q tile, then coalesced 16-byte stores
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld<32>(trow + kFwdColO + half * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 xa = *reinterpret_cast<const float4*>(vx + half * 32 + g * 8);
                const float4 xb = *reinterpret_cast<const float4*>(vx + half * 32 + g * 8 + 4);
                const uint4 o = make_uint4(
                    pack_bf16(fmaf(px, xa.x, __uint_as_float(v[8 * g])) * inv,
                              fmaf(px, xa.y, __uint_as_float(v[8 * g + 1])) * inv),
                    pack_bf16(fmaf(px, xa.z, __uint_as_float(v[8 * g + 2])) * inv,
                              fmaf(px, xa.w, __uint_as_float(v[8 * g + 3])) * inv),
                    pack_bf16(fmaf(px, xb.x, __uint_as_float(v[8 * g + 4])) * inv,
                              fmaf(px, xb.y, __uint_as_float(v[8 * g + 5])) * inv),
                    pack_bf16(fmaf(px, xb.z, __uint_as_float(v[8 * g + 6])) * inv,
                              fmaf(px, xb.w, __uint_as_float(v[8 * g + 7])) * inv));
                *reinterpret_cast<uint4*>(sm_q + row_chunk(r, half * 4 + g)) = o;
            }
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
    } else if (has_edge_row) {
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
        }
        const float exx = exp2f(fmaf(sxx, kLog2e, -mb));
        const float sum = warp_sum(part) + exx;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) pbuf[lane + 32 * jj] = s[jj];
        __syncwarp();
        mbar_wait(&bars[1], 0);
        edge_gemv(pbuf, sm_v, nv, exx, vx, p.out + (static_cast<size_t>(n) * p.T + nv) * D + h * kHd, 1.0f / sum, lane);
        if (lane == 0) p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + nv] = mx + logf(sum);
        PCG_TRACE(9);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) PCG_TRACE(10);
    if (warp == 4) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 384;  // warps 0-7 elementwise (lane quarter = warp & 3, column half = warp >> 2),
                                  // warp 8 TMA + MMA, warps 9-11 edge rows (dV_x, dK_x, dQ_x)
constexpr int kBwdOffQ = 0;
constexpr int kBwdOffK = 2 * kBlkBytes;
constexpr int kBwdOffV = 4 * kBlkBytes;
constexpr int kBwdOffDO = 6 * kBlkBytes;
constexpr int kBwdOffPT = 8 * kBlkBytes;    // P^T  [128 keys x 128 queries] bf16, two 64-column blocks
constexpr int kBwdOffDST = 10 * kBlkBytes;  // dS^T, same layout
constexpr int kBwdOffStage = 12 * kBlkBytes;  // 8 warps x [32 rows x 64 B] epilogue staging
constexpr int kBwdOffVec = 13 * kBlkBytes;  // float[256] x 6: lse2, delta, pcol, dscol, prow, dsrow
constexpr int kBwdOffX = kBwdOffVec + 6 * 1024;  // float[64] x 4: q_x, k_x, v_x, dO_x; then 8 scalars
constexpr int kBwdOffBar = kBwdOffX + 4 * 256 + 32;
constexpr int kBwdSmemBytes = kBwdOffBar + 64 + 1024;
constexpr uint32_t kColST = 0, kColDPT = 128, kColDV = 256, kColDK = 320, kColDQ = 384;

struct BwdParams {
    int T, heads;
    int nv;       // T - 1
    int n_tiles;  // ceil(nv / 128): 1 or 2
    const bf16* qkv;
    const bf16* d_out;
    const float* lse;
    const float* delta;
    bf16* d_qkv;
    long long* trace;
};

// W columns of one block: P^T and dS^T for this thread's key row, written as bf16 into the swizzled smem blocks
template <int W>
__device__ __forceinline__ void bwd_chunk(uint32_t t_s, uint32_t t_dp, int c, const float* lse2, const float* delta,
                                          bool row_ok, uint8_t* pt_blk, uint8_t* dst_blk, int r) {
    uint32_t s[W], dp[W];
    tmem_ld<W>(t_s + c, s);
    tmem_ld<W>(t_dp + c, dp);
    tmem_wait_ld();
#pragma unroll
    for (int g = 0; g < W / 8; ++g) {
        const float4 la = *reinterpret_cast<const float4*>(lse2 + c + 8 * g);
        const float4 lb = *reinterpret_cast<const float4*>(lse2 + c + 8 * g + 4);
        const float4 da = *reinterpret_cast<const float4*>(delta + c + 8 * g);
        const float4 db = *reinterpret_cast<const float4*>(delta + c + 8 * g + 4);
        const float l[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
        const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        float pv[8], dsv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            pv[k] = exp2f(fmaf(__uint_as_float(s[8 * g + k]), kLog2e, -l[k]));
            dsv[k] = pv[k] * (__uint_as_float(dp[8 * g + k]) - d[k]);
        }
        uint4 pq = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                              pack_bf16(pv[6], pv[7]));
        uint4 dq = make_uint4(pack_bf16(dsv[0], dsv[1]), pack_bf16(dsv[2], dsv[3]), pack_bf16(dsv[4], dsv[5]),
                              pack_bf16(dsv[6], dsv[7]));
        if (!row_ok) pq = dq = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t off = row_chunk(r, ((c & 63) >> 3) + g);
        *reinterpret_cast<uint4*>(pt_blk + off) = pq;
        *reinterpret_cast<uint4*>(dst_blk + off) = dq;
    }
}

// 32 accumulator columns of this thread's row + coef * xrow[.] -> bf16 -> the warp's staging rows -> global
__device__ __forceinline__ void bwd_epilogue(uint32_t taddr, float coef, const float* xrow, uint8_t* stage, int lane,
                                             bf16* gbase, size_t ld, int row0, int row_end) {
    uint32_t v[32];
    tmem_ld<32>(taddr, v);
    tmem_wait_ld();
    const int sw = (lane >> 1) & 3;
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

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                   const BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sm_q = sm + kBwdOffQ;
    uint8_t* sm_k = sm + kBwdOffK;
    uint8_t* sm_v = sm + kBwdOffV;
    uint8_t* sm_do = sm + kBwdOffDO;
    uint8_t* sm_pt = sm + kBwdOffPT;
    uint8_t* sm_dst = sm + kBwdOffDST;
    float* lse2 = reinterpret_cast<float*>(sm + kBwdOffVec);
    float* delta = lse2 + 256;
    float* pcol = lse2 + 512;   // p(query i, key x)
    float* dscol = lse2 + 768;  // ds(query i, key x)
    float* prow = lse2 + 1024;  // p(query x, key r)
    float* dsrow = lse2 + 1280; // ds(query x, key r)
    float* qx = reinterpret_cast<float*>(sm + kBwdOffX);
    float* kx = qx + 64;
    float* vx = qx + 128;
    float* dox = qx + 192;
    float* scal = qx + 256;  // [0] p_xx, [1] ds_xx
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kBwdOffBar);  // 0 loads, 1 S^T/dP^T, 2 P^T/dS^T, 3 tile done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, n = blockIdx.y;
    const int D = p.heads * kHd, nv = p.nv, nt = p.n_tiles;
    const size_t vbase = (static_cast<size_t>(n) * p.heads + h) * p.T;
    const float lse2_x = p.lse[vbase + nv] * kLog2e, delta_x = p.delta[vbase + nv];
    const size_t cta_id = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
    if (warp == 0) PCG_TRACE(0);

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            mbar_init(&bars[2], 8);
            mbar_init(&bars[3], 1);
            fence_barrier_init();
            tma_prefetch_desc(&map_qkv);
            tma_prefetch_desc(&map_do);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    } else if (warp < 8) {
        const int t = threadIdx.x;
        lse2[t] = (t < nv) ? p.lse[vbase + t] * kLog2e : INFINITY;
        delta[t] = (t < nv) ? p.delta[vbase + t] : 0.f;
    } else {
        const bf16* xq = p.qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + h * kHd;
        if (warp == 9) load_row_f32(qx, xq, lane), load_row_f32(dox, p.d_out + (static_cast<size_t>(n) * p.T + nv) * D + h * kHd, lane);
        if (warp == 10) load_row_f32(kx, xq + D, lane);
        if (warp == 11) load_row_f32(vx, xq + 2 * D, lane);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars[0], 4 * nt * kBlkBytes);
            for (int i = 0; i < nt; ++i) {
                tma_load_3d(&map_qkv, &bars[0], sm_k + i * kBlkBytes, D + h * kHd, i * 128, n, kEvictFirst);
                tma_load_3d(&map_qkv, &bars[0], sm_q + i * kBlkBytes, h * kHd, i * 128, n, kEvictFirst);
                tma_load_3d(&map_qkv, &bars[0], sm_v + i * kBlkBytes, 2 * D + h * kHd, i * 128, n, kEvictFirst);
                tma_load_3d(&map_do, &bars[0], sm_do + i * kBlkBytes, h * kHd, i * 128, n, kEvictFirst);
            }
            mbar_wait(&bars[0], 0);
            PCG_TRACE(1);
            tc_fence_after();
            const int n_blocks = nt * nt;
            // block b = (key tile j, query block i), j-major; widths of the query blocks
            auto width = [&](int i) { return min(128, ((nv - 128 * i) + 15) & ~15); };
            mma_tile_x_rows(tmem + kColST, sm_k, sm_q, width(0));    // S^T  = K_0 Q_0^T
            mma_tile_x_rows(tmem + kColDPT, sm_v, sm_do, width(0));  // dP^T = V_0 dO_0^T
            umma_commit(&bars[1]);
            for (int b = 0; b < n_blocks; ++b) {
                const int j = b / nt, i = b - j * nt;
                const int ksteps = width(i) >> 4;
                mbar_wait(&bars[2], b & 1);  // P^T and dS^T of block b are in shared memory, S^T / dP^T consumed
                PCG_TRACE(16 + b);
                tc_fence_after();
                mma_blocks_x_cols(tmem + kColDV, sm_pt, sm_do + i * kBlkBytes, ksteps, i != 0);   // dV_j += P^T dO_i
                mma_blocks_x_cols(tmem + kColDK, sm_dst, sm_q + i * kBlkBytes, ksteps, i != 0);   // dK_j += dS^T Q_i
                mma_rows_t_x_cols(tmem + kColDQ + 64 * i, sm_dst, sm_k + j * kBlkBytes, 8, j != 0);  // dQ_i += dS K_j
                if (i == nt - 1) umma_commit(&bars[3]);  // dV_j, dK_j complete (and dQ after the last tile)
                if (b + 1 < n_blocks) {
                    const int j2 = (b + 1) / nt, i2 = (b + 1) - j2 * nt;
                    mma_tile_x_rows(tmem + kColST, sm_k + j2 * kBlkBytes, sm_q + i2 * kBlkBytes, width(i2));
                    mma_tile_x_rows(tmem + kColDPT, sm_v + j2 * kBlkBytes, sm_do + i2 * kBlkBytes, width(i2));
                    umma_commit(&bars[1]);
                }
            }
        }
    } else {
        mbar_wait(&bars[0], 0);  // Q, K, V, dO are in shared memory
        if (warp < 8) {
            // edge products that need one dot product per token: thread t = query t (column x of S) and key t (row x)
            const int t = threadIdx.x;
            if (t < nv) {
                const float pc = exp2f(fmaf(row_dot(sm_q, t, kx), kLog2e, -lse2[t]));
                pcol[t] = pc;
                dscol[t] = pc * (row_dot(sm_do, t, vx) - delta[t]);
                const float pr = exp2f(fmaf(row_dot(sm_k, t, qx), kLog2e, -lse2_x));
                prow[t] = pr;
                dsrow[t] = pr * (row_dot(sm_v, t, dox) - delta_x);
            } else {
                pcol[t] = dscol[t] = prow[t] = dsrow[t] = 0.f;
            }
            if (t == 0) {
                float sxx = 0.f, dpxx = 0.f;
                for (int d = 0; d < 64; ++d) sxx = fmaf(qx[d], kx[d], sxx), dpxx = fmaf(dox[d], vx[d], dpxx);
                const float pxx = exp2f(fmaf(sxx, kLog2e, -lse2_x));
                scal[0] = pxx;
                scal[1] = pxx * (dpxx - delta_x);
            }
        }
        named_bar_sync(1, kBwdThreads - 32);  // edge vectors visible to the elementwise and the edge warps
        if (warp == 0) PCG_TRACE(2);
        if (warp >= 9) {
            bf16* gx = p.d_qkv + (static_cast<size_t>(n) * p.T + nv) * 3 * D + h * kHd;
            if (warp == 9) edge_gemv(pcol, sm_do, nv, scal[0], dox, gx + 2 * D, 1.0f, lane);  // dV_x
            if (warp == 10) edge_gemv(dscol, sm_q, nv, scal[1], qx, gx + D, 1.0f, lane);       // dK_x
            if (warp == 11) edge_gemv(dsrow, sm_k, nv, scal[1], kx, gx, 1.0f, lane);           // dQ_x
            if (warp == 11) PCG_TRACE(3);
        } else {
            const int quarter = warp & 3, half = warp >> 2;
            const int r = quarter * 32 + lane;  // key row in the tile == TMEM lane
            const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
            uint8_t* stage = sm + kBwdOffStage + warp * 2048;
            bf16* gd = p.d_qkv + static_cast<size_t>(n) * p.T * 3 * D + h * kHd + half * 32;
            int b = 0;
            for (int j = 0; j < nt; ++j) {
                const bool row_ok = (j * 128 + r) < nv;
                for (int i = 0; i < nt; ++i, ++b) {
                    const int width = min(128, ((nv - 128 * i) + 15) & ~15);
                    mbar_wait(&bars[1], b & 1);
                    tc_fence_after();
                    if (warp == 0) PCG_TRACE(4 + 2 * b);
                    uint8_t* pt_blk = sm_pt + half * kBlkBytes;
                    uint8_t* dst_blk = sm_dst + half * kBlkBytes;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c = half * 64 + cc * 32;
                        if (c + 32 <= width)
                            bwd_chunk<32>(trow + kColST, trow + kColDPT, c, lse2 + i * 128, delta + i * 128, row_ok,
                                          pt_blk, dst_blk, r);
                        else if (c < width)
                            bwd_chunk<16>(trow + kColST, trow + kColDPT, c, lse2 + i * 128, delta + i * 128, row_ok,
                                          pt_blk, dst_blk, r);
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[2]);
                    if (warp == 0) PCG_TRACE(5 + 2 * b);
                }
                // key tile j finished: dV_j and dK_j (+ the edge query's contribution) -> global
                mbar_wait(&bars[3], j & 1);
                tc_fence_after();
                if (warp == 0) PCG_TRACE(12 + j);
                const int row0 = j * 128 + quarter * 32;
                bwd_epilogue(trow + kColDV + half * 32, prow[j * 128 + r], dox + half * 32, stage, lane, gd + 2 * D,
                             static_cast<size_t>(3) * D, row0, nv);
                bwd_epilogue(trow + kColDK + half * 32, dsrow[j * 128 + r], qx + half * 32, stage, lane, gd + D,
                             static_cast<size_t>(3) * D, row0, nv);
                tc_fence_before();
            }
            // dQ_i (+ the edge key's contribution)
            for (int i = 0; i < nt; ++i)
                bwd_epilogue(trow + kColDQ + 64 * i + half * 32, dscol[i * 128 + r], kx + half * 32, stage, lane, gd,
                             static_cast<size_t>(3) * D, i * 128 + quarter * 32, nv);
        }
    }
    if (warp == 0) PCG_TRACE(14);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) PCG_TRACE(15);
    if (warp == 8) tmem_dealloc(tmem, 512);
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

// [n][T][cols] bf16 view of a row-major [n*T, cols] matrix; box = {64 columns, 128 rows, 1}.  Rows >= T of a box
// are zero-filled, so whole 128-row boxes are always loaded.
int make_map3(CUtensorMap* out, const void* ptr, int n, int T, int cols) {
    struct Key {
        const void* p;
        int n, T, cols;
        bool operator==(const Key& o) const { return p == o.p && n == o.n && T == o.T && cols == o.cols; }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            size_t h = reinterpret_cast<size_t>(k.p);
            for (int v : {k.n, k.T, k.cols}) h = h * 1000003u ^ static_cast<size_t>(v);
            return h;
        }
    };
    static std::mutex mu;
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    const Key key{ptr, n, T, cols};
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
    const cuuint32_t box[3] = {64, 128, 1};
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

// tensor-core path: T - 1 tokens in one or two 128-row tiles, at least 65 so the tile is not mostly padding
bool use_tc(int T) { return T >= 66 && T <= 257; }

long long* g_trace = nullptr;

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
    if (g_force_legacy || !use_tc(T)) return attn_fwd_legacy(qkv, out, lse, n, T, heads, 0, s);
    const int D = heads * kHd;
    CUtensorMap map;
    if (int rc = make_map3(&map, qkv, n, T, 3 * D)) return rc;
    static bool configured = false;
    if (!configured) {
        PCG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
        configured = true;
    }
    const int nv = T - 1;
    FwdParams p{T, heads, nv, (nv + 15) & ~15, static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse,
                g_trace};
    attn_fwd_tc_kernel<<<dim3((nv + 127) / 128, heads, n), kFwdThreads, kFwdSmemBytes, s>>>(map, p);
    PCG_LAUNCH_CHECK("attn_fwd_tc_kernel");
    return 0;
}

extern "C" int pcg_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                            void* d_qkv, int n, int T, int heads, void* stream) {
    PCG_CHECK_ARG(qkv && out && d_out && lse && delta_ws && d_qkv, "pcg_attn_bwd: null pointer");
    PCG_CHECK_ARG(n > 0 && T > 0 && heads > 0, "pcg_attn_bwd: bad shape n=%d T=%d heads=%d", n, T, heads);
    PCG_CHECK_ARG(n <= 65535 && heads <= 65535, "pcg_attn_bwd: n and heads must be <= 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_BWD, 8.0 * T * T * kHd * heads * n, s);
    if (int rc = attn_delta(out, d_out, delta_ws, n, T, heads, s)) return rc;
    if (g_force_legacy || !use_tc(T)) return attn_bwd_legacy(qkv, d_out, lse, delta_ws, d_qkv, n, T, heads, 0, s);
    const int D = heads * kHd;
    CUtensorMap map, map_do;
    if (int rc = make_map3(&map, qkv, n, T, 3 * D)) return rc;
    if (int rc = make_map3(&map_do, d_out, n, T, D)) return rc;
    static bool configured = false;
    if (!configured) {
        PCG_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
        configured = true;
    }
    const int nv = T - 1;
    BwdParams p{T, heads, nv, (nv + 127) / 128, static_cast<const bf16*>(qkv), static_cast<const bf16*>(d_out),
                lse, delta_ws, static_cast<bf16*>(d_qkv), g_trace};
    attn_bwd_tc_kernel<<<dim3(heads, n), kBwdThreads, kBwdSmemBytes, s>>>(map, map_do, p);
    PCG_LAUNCH_CHECK("attn_bwd_tc_kernel");
    return 0;
}
