// Fused short-sequence multi-head attention (T in {50,197,257,577}, head dim 64, no mask), forward and backward.
//
// One CTA = one (cutout, head, 64-row tile); 4 warps x 16 rows.  Q/K/V tiles are staged with cp.async into
// XOR-swizzled shared memory (double buffered), S = QK^T and P.V run on mma.sync m16n8k16 bf16 with the
// softmax kept in registers (flash style: P never touches memory).  The backward is two deterministic passes
// (dK/dV over key tiles, dQ over query tiles) that recompute P from Q, K and the saved log-sum-exp.
// Attention is 1.6-12 % of the step FLOPs; the GEMMs around it are the tcgen05 kernels.
//
// Layout: qkv bf16 [n*T, 3D], row = cutout*T + token, columns [q | k | v], head h at columns h*64..h*64+63 of
// each third; q is pre-scaled by 1/8 (folded into the weights).  Replaces nn.MultiheadAttention's core in
// perceptor/models/ruclip/model.py:43-49.
#include "pcg_common.cuh"
#include "pcg_ptx.cuh"

namespace pcg {
namespace {

constexpr int kTile = 64;  // rows per tile (queries or keys)
constexpr int kHd = 64;    // head dim
constexpr float kLog2e = 1.4426950408889634f;

using bf16 = __nv_bfloat16;

// byte offset of element (row, col) in a swizzled [64][64] bf16 tile: 16-byte chunk index XOR (row & 7)
__device__ __forceinline__ uint32_t swz(int row, int col) {
    return static_cast<uint32_t>(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}

// cp.async a [64 x 64] bf16 tile: rows t0..t0+63 of a [T, ld] matrix (rows >= T are zero-filled).
__device__ __forceinline__ void load_tile_async(bf16* stile, const bf16* gbase, int ld, int t0, int T) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int id = threadIdx.x + i * 128;
        const int row = id >> 3, ch = id & 7;
        const bool ok = (t0 + row) < T;
        const bf16* src = gbase + static_cast<size_t>(ok ? (t0 + row) : 0) * ld + ch * 8;
        cp_async_16(reinterpret_cast<uint8_t*>(stile) + row * 128 + ((ch ^ (row & 7)) << 4), src, ok);
    }
}

// A fragments (16 rows x 64 k) of rows [row0, row0+16) of a swizzled tile: frag[ks][0..3]
__device__ __forceinline__ void load_a_frags(uint32_t (&frag)[4][4], const bf16* stile, int row0) {
    const int lane = threadIdx.x & 31;
    const uint32_t base = smem_u32(stile);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int row = row0 + (lane & 15);
        const int col = ks * 16 + ((lane >> 4) << 3);
        ldmatrix_x4(frag[ks], base + swz(row, col));
    }
}

// B fragments for C[m, n] += A[m, k] * Tile[n, k]   (tile rows index n, tile columns index k; "non-transposed").
// Returns the fragments of n-tiles (2*pair, 2*pair+1) for k-step ks: {b0,b1} and {b2,b3}.
__device__ __forceinline__ void load_b_nk(uint32_t (&r)[4], const bf16* stile, int pair, int ks) {
    const int lane = threadIdx.x & 31;
    const int mi = lane >> 3;
    const int row = pair * 16 + ((mi >> 1) << 3) + (lane & 7);
    const int col = ks * 16 + ((mi & 1) << 3);
    ldmatrix_x4(r, smem_u32(stile) + swz(row, col));
}
// B fragments for C[m, n] += A[m, k] * Tile[k, n]   (tile rows index k; needs the transposing ldmatrix).
// Returns fragments of n-tiles (2*pair, 2*pair+1) for k-step ks.
__device__ __forceinline__ void load_b_kn(uint32_t (&r)[4], const bf16* stile, int pair, int ks) {
    const int lane = threadIdx.x & 31;
    const int mi = lane >> 3;
    const int row = ks * 16 + ((mi & 1) << 3) + (lane & 7);
    const int col = pair * 16 + ((mi >> 1) << 3);
    ldmatrix_x4_trans(r, smem_u32(stile) + swz(row, col));
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Write a warp's 16 x 64 fp32 accumulator tile as bf16 through its own 16 rows of a swizzled smem tile, then
// to global with 16-byte stores.  acc[nt][0..3]: n-tile nt (8 columns), c0,c1 row lane/4, c2,c3 row lane/4+8.
__device__ __forceinline__ void store_tile_bf16(const float (&acc)[8][4], bf16* stile, int row0, bf16* gbase, int ld,
                                                int t0, int T) {
    const int lane = threadIdx.x & 31;
    uint8_t* sb = reinterpret_cast<uint8_t*>(stile);
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + ((lane & 3) << 1);
        const int r0 = row0 + (lane >> 2);
        *reinterpret_cast<uint32_t*>(sb + swz(r0, col)) = pack_bf16(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<uint32_t*>(sb + swz(r0 + 8, col)) = pack_bf16(acc[nt][2], acc[nt][3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int id = lane + i * 32;
        const int row = row0 + (id >> 3), ch = id & 7;
        if (t0 + row < T) {
            const uint4 v = *reinterpret_cast<const uint4*>(sb + row * 128 + ((ch ^ (row & 7)) << 4));
            *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(t0 + row) * ld + ch * 8) = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
// NB = 64-column blocks per head: 1 for head dim 64; 2 for the wide heads of ViT-H/14 (80) and ViT-g/14 (88), which
// the weight packing pads to 128 columns with zero rows (perceptor_b200/vit.py), so S = Q K^T simply runs over two
// blocks of k-steps and O carries two 64-column halves.  Shared-memory tiles stay [64 x 64] swizzled blocks.
template <int NB>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                       float* __restrict__ lse, int T, int heads, int q_begin) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    constexpr int HD = kHd * NB;
    constexpr int kBlk = kTile * kHd;  // elements of one [64 x 64] block
    extern __shared__ __align__(128) uint8_t smem_fwd[];
    bf16* sQ = reinterpret_cast<bf16*>(smem_fwd);      // [NB][kBlk]
    bf16* sK = sQ + NB * kBlk;                          // [2][NB][kBlk]
    bf16* sV = sK + 2 * NB * kBlk;                      // [2][NB][kBlk]

    const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int D = heads * HD, ld = 3 * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = q_begin + qt * kTile;  // rows below q_begin belong to the tcgen05 kernel
    const bf16* gq = qkv + static_cast<size_t>(n) * T * ld + h * HD;
    const bf16* gk = gq + D;
    const bf16* gv = gq + 2 * D;
    const int nkv = (T + kTile - 1) / kTile;

#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
        load_tile_async(sQ + blk * kBlk, gq + blk * kHd, ld, q0, T);
        load_tile_async(sK + blk * kBlk, gk + blk * kHd, ld, 0, T);
        load_tile_async(sV + blk * kBlk, gv + blk * kHd, ld, 0, T);
    }
    cp_async_commit();

    const bool warp_active = (q0 + warp * 16) < T;  // warp-uniform
    uint32_t qf[NB][4][4];
    float o[NB * 8][4];
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NB * 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

    for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j + 1 < nkv) {
#pragma unroll
            for (int blk = 0; blk < NB; ++blk) {
                load_tile_async(sK + ((buf ^ 1) * NB + blk) * kBlk, gk + blk * kHd, ld, (j + 1) * kTile, T);
                load_tile_async(sV + ((buf ^ 1) * NB + blk) * kBlk, gv + blk * kHd, ld, (j + 1) * kTile, T);
            }
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            if (j == 0) {
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) load_a_frags(qf[blk], sQ + blk * kBlk, warp * 16);
            }
            const int valid = min(kTile, T - j * kTile);  // valid keys in this tile
            const int npair = (valid + 15) >> 4;
            float s[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) s[i][c] = 0.f;
#pragma unroll
            for (int pair = 0; pair < 4; ++pair) {
                if (pair < npair) {
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            uint32_t b[4];
                            load_b_nk(b, sK + (buf * NB + blk) * kBlk, pair, ks);
                            mma_bf16_16816(s[2 * pair], qf[blk][ks], b[0], b[1]);
                            mma_bf16_16816(s[2 * pair + 1], qf[blk][ks], b[2], b[3]);
                        }
                    }
                }
            }
            // mask + running max
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int col = nt * 8 + ((lane & 3) << 1);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool ok = (col + (c & 1)) < valid;
                    s[nt][c] = ok ? s[nt][c] : -INFINITY;
                    mx[c >> 1] = fmaxf(mx[c >> 1], s[nt][c]);
                }
            }
            float alpha[2], mnew[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                mnew[r] = fmaxf(m_run[r], quad_max(mx[r]));
                alpha[r] = exp2f((m_run[r] - mnew[r]) * kLog2e);
                m_run[r] = mnew[r];
            }
            float rs[2] = {0.f, 0.f};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float pv = exp2f((s[nt][c] - mnew[c >> 1]) * kLog2e);
                    s[nt][c] = pv;
                    rs[c >> 1] += pv;
                }
#pragma unroll
            for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + quad_sum(rs[r]);
#pragma unroll
            for (int nt = 0; nt < NB * 8; ++nt) {
                o[nt][0] *= alpha[0];
                o[nt][1] *= alpha[0];
                o[nt][2] *= alpha[1];
                o[nt][3] *= alpha[1];
            }
            // O += P V
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (ks < npair) {
                    uint32_t pa[4];
                    pa[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
                    pa[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
                    pa[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
                    pa[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
#pragma unroll
                        for (int dp = 0; dp < 4; ++dp) {
                            uint32_t b[4];
                            load_b_kn(b, sV + (buf * NB + blk) * kBlk, dp, ks);
                            mma_bf16_16816(o[blk * 8 + 2 * dp], pa, b[0], b[1]);
                            mma_bf16_16816(o[blk * 8 + 2 * dp + 1], pa, b[2], b[3]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone done with buf before it is refilled two iterations later
    }

    if (warp_active) {
        const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
        for (int nt = 0; nt < NB * 8; ++nt) {
            o[nt][0] *= inv0;
            o[nt][1] *= inv0;
            o[nt][2] *= inv1;
            o[nt][3] *= inv1;
        }
        bf16* go = out + static_cast<size_t>(n) * T * D + h * HD;
#pragma unroll
        for (int blk = 0; blk < NB; ++blk)
            store_tile_bf16(reinterpret_cast<const float(&)[8][4]>(o[blk * 8]), sQ + blk * kBlk, warp * 16, go + blk * kHd, D,
                            q0, T);
        if ((lane & 3) == 0) {
            const int r0 = q0 + warp * 16 + (lane >> 2);
            float* gl = lse + (static_cast<size_t>(n) * heads + h) * T;
            if (r0 < T) gl[r0] = m_run[0] + logf(l_run[0]);
            if (r0 + 8 < T) gl[r0 + 8] = m_run[1] + logf(l_run[1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward, pass 0: delta[n,h,t] = sum_d dO * O
// ---------------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                         float* __restrict__ delta, int n_rows, int T, int heads) {
    // 8 * NB lanes x 16 bytes cover the 64 * NB dims of one (row, head); a warp handles 4 / NB consecutive pairs
    constexpr int kLanes = 8 * NB, HD = kHd * NB;
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long pair = gid / kLanes;
    const int sub = threadIdx.x & (kLanes - 1);
    const bool ok = pair < static_cast<long long>(n_rows) * heads;
    float v = 0.f;
    int row = 0, h = 0;
    if (ok) {
        row = static_cast<int>(pair / heads);
        h = static_cast<int>(pair % heads);
        const size_t off = static_cast<size_t>(row) * heads * HD + h * HD + sub * 8;
        const uint4 a = *reinterpret_cast<const uint4*>(o + off);
        const uint4 b = *reinterpret_cast<const uint4*>(d_o + off);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&aw[q]);
            const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(&bw[q]);
            v = fmaf(__low2float(x), __low2float(y), v);
            v = fmaf(__high2float(x), __high2float(y), v);
        }
    }
#pragma unroll
    for (int m = 1; m < kLanes; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (ok && sub == 0) {
        const int nn = row / T, t = row % T;
        delta[(static_cast<size_t>(nn) * heads + h) * T + t] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward, pass 1: dK, dV.  CTA = 64 keys; loops over query tiles.  Works on the transposed problem
// S^T = K Q^T so that P^T / dS^T land directly in A-fragment layout for the dV / dK products.
// ---------------------------------------------------------------------------------------------------------
template <int NB>
struct BwdSmem {
    bf16 kv[2][NB][kTile * kHd];  // K and V tiles of this CTA (only read once into registers)
    bf16 q[2][NB][kTile * kHd];
    bf16 d_o[2][NB][kTile * kHd];
    float lse[2][kTile];
    float delta[2][kTile];
};

__device__ __forceinline__ void load_vec_async(float* sdst, const float* gsrc, int t0, int T, float fill) {
    if (threadIdx.x < kTile) {
        const int t = t0 + threadIdx.x;
        sdst[threadIdx.x] = (t < T) ? gsrc[t] : fill;
    }
}

template <int NB>
__global__ void __launch_bounds__(128, NB == 1 ? 3 : 1)
attn_bwd_dkdv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_out, const float* __restrict__ lse,
                     const float* __restrict__ delta, bf16* __restrict__ d_qkv, int T, int heads, int k_begin) {
    constexpr int HD = kHd * NB;
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    BwdSmem<NB>& sm = *reinterpret_cast<BwdSmem<NB>*>(smem_dyn);

    const int kt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int D = heads * HD, ld = 3 * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = k_begin + kt * kTile;  // keys below k_begin belong to the tcgen05 kernel
    const bf16* gq = qkv + static_cast<size_t>(n) * T * ld + h * HD;
    const bf16* gk = gq + D;
    const bf16* gv = gq + 2 * D;
    const bf16* gdo = d_out + static_cast<size_t>(n) * T * D + h * HD;
    const float* glse = lse + (static_cast<size_t>(n) * heads + h) * T;
    const float* gdel = delta + (static_cast<size_t>(n) * heads + h) * T;
    const int nq = (T + kTile - 1) / kTile;

#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
        load_tile_async(sm.kv[0][blk], gk + blk * kHd, ld, k0, T);
        load_tile_async(sm.kv[1][blk], gv + blk * kHd, ld, k0, T);
        load_tile_async(sm.q[0][blk], gq + blk * kHd, ld, 0, T);
        load_tile_async(sm.d_o[0][blk], gdo + blk * kHd, D, 0, T);
    }
    cp_async_commit();
    load_vec_async(sm.lse[0], glse, 0, T, INFINITY);
    load_vec_async(sm.delta[0], gdel, 0, T, 0.f);

    const bool warp_active = (k0 + warp * 16) < T;
    float dk[NB * 8][4], dv[NB * 8][4];
#pragma unroll
    for (int i = 0; i < NB * 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;

    for (int i = 0; i < nq; ++i) {
        const int buf = i & 1;
        if (i + 1 < nq) {
#pragma unroll
            for (int blk = 0; blk < NB; ++blk) {
                load_tile_async(sm.q[buf ^ 1][blk], gq + blk * kHd, ld, (i + 1) * kTile, T);
                load_tile_async(sm.d_o[buf ^ 1][blk], gdo + blk * kHd, D, (i + 1) * kTile, T);
            }
            cp_async_commit();
            load_vec_async(sm.lse[buf ^ 1], glse, (i + 1) * kTile, T, INFINITY);
            load_vec_async(sm.delta[buf ^ 1], gdel, (i + 1) * kTile, T, 0.f);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            const int valid = min(kTile, T - i * kTile);  // valid queries in this tile
            const int npair = (valid + 15) >> 4;
            // two halves of 32 query columns each: halves the live S^T / dP^T registers (3 CTAs per SM instead of 2)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (2 * half < npair) {
                    float st[4][4], dpt[4][4];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int c = 0; c < 4; ++c) st[a][c] = dpt[a][c] = 0.f;
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            // A fragments of this warp's 16 keys are re-read from shared memory instead of being pinned
                            uint32_t ka[4], va[4];
                            {
                                const int row = warp * 16 + (lane & 15), col = ks * 16 + ((lane >> 4) << 3);
                                ldmatrix_x4(ka, smem_u32(sm.kv[0][blk]) + swz(row, col));
                                ldmatrix_x4(va, smem_u32(sm.kv[1][blk]) + swz(row, col));
                            }
#pragma unroll
                            for (int pp = 0; pp < 2; ++pp) {
                                const int pair = 2 * half + pp;
                                if (pair < npair) {
                                    uint32_t b[4];
                                    load_b_nk(b, sm.q[buf][blk], pair, ks);
                                    mma_bf16_16816(st[2 * pp], ka, b[0], b[1]);
                                    mma_bf16_16816(st[2 * pp + 1], ka, b[2], b[3]);
                                    load_b_nk(b, sm.d_o[buf][blk], pair, ks);
                                    mma_bf16_16816(dpt[2 * pp], va, b[0], b[1]);
                                    mma_bf16_16816(dpt[2 * pp + 1], va, b[2], b[3]);
                                }
                            }
                        }
                    }
                    // P^T and dS^T (columns are queries)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int col = half * 32 + nt * 8 + ((lane & 3) << 1);
                        const float l0 = sm.lse[buf][col], l1 = sm.lse[buf][col + 1];
                        const float d0 = sm.delta[buf][col], d1 = sm.delta[buf][col + 1];
                        const float p0 = exp2f((st[nt][0] - l0) * kLog2e), p1 = exp2f((st[nt][1] - l1) * kLog2e);
                        const float p2 = exp2f((st[nt][2] - l0) * kLog2e), p3 = exp2f((st[nt][3] - l1) * kLog2e);
                        st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
                        dpt[nt][0] = p0 * (dpt[nt][0] - d0);
                        dpt[nt][1] = p1 * (dpt[nt][1] - d1);
                        dpt[nt][2] = p2 * (dpt[nt][2] - d0);
                        dpt[nt][3] = p3 * (dpt[nt][3] - d1);
                    }
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        const int ks = 2 * half + pp;  // k-step over queries
                        if (ks < npair) {
                            uint32_t pa[4], da[4];
                            pa[0] = pack_bf16(st[2 * pp][0], st[2 * pp][1]);
                            pa[1] = pack_bf16(st[2 * pp][2], st[2 * pp][3]);
                            pa[2] = pack_bf16(st[2 * pp + 1][0], st[2 * pp + 1][1]);
                            pa[3] = pack_bf16(st[2 * pp + 1][2], st[2 * pp + 1][3]);
                            da[0] = pack_bf16(dpt[2 * pp][0], dpt[2 * pp][1]);
                            da[1] = pack_bf16(dpt[2 * pp][2], dpt[2 * pp][3]);
                            da[2] = pack_bf16(dpt[2 * pp + 1][0], dpt[2 * pp + 1][1]);
                            da[3] = pack_bf16(dpt[2 * pp + 1][2], dpt[2 * pp + 1][3]);
#pragma unroll
                            for (int blk = 0; blk < NB; ++blk) {
#pragma unroll
                                for (int dp = 0; dp < 4; ++dp) {
                                    uint32_t b[4];
                                    load_b_kn(b, sm.d_o[buf][blk], dp, ks);
                                    mma_bf16_16816(dv[blk * 8 + 2 * dp], pa, b[0], b[1]);
                                    mma_bf16_16816(dv[blk * 8 + 2 * dp + 1], pa, b[2], b[3]);
                                    load_b_kn(b, sm.q[buf][blk], dp, ks);
                                    mma_bf16_16816(dk[blk * 8 + 2 * dp], da, b[0], b[1]);
                                    mma_bf16_16816(dk[blk * 8 + 2 * dp + 1], da, b[2], b[3]);
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    if (warp_active) {
        bf16* gdk = d_qkv + static_cast<size_t>(n) * T * ld + D + h * HD;
        bf16* gdv = gdk + D;
        // each warp reuses its own 16 rows of the (now dead) K / V staging tiles
#pragma unroll
        for (int blk = 0; blk < NB; ++blk) {
            store_tile_bf16(reinterpret_cast<const float(&)[8][4]>(dk[blk * 8]), sm.kv[0][blk], warp * 16, gdk + blk * kHd, ld,
                            k0, T);
            store_tile_bf16(reinterpret_cast<const float(&)[8][4]>(dv[blk * 8]), sm.kv[1][blk], warp * 16, gdv + blk * kHd, ld,
                            k0, T);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward, pass 2: dQ.  CTA = 64 queries; loops over key tiles.
// ---------------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_out,
                                                          const float* __restrict__ lse,
                                                          const float* __restrict__ delta, bf16* __restrict__ d_qkv,
                                                          int T, int heads, int q_begin) {
    grid_dep_launch();  // a dependent (PDL) kernel may start its prologue while this grid drains
    constexpr int HD = kHd * NB;
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    BwdSmem<NB>& sm = *reinterpret_cast<BwdSmem<NB>*>(smem_dyn);  // kv[] holds Q / dO here; q[] / d_o[] hold K / V tiles

    const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int D = heads * HD, ld = 3 * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = q_begin + qt * kTile;
    const bf16* gq = qkv + static_cast<size_t>(n) * T * ld + h * HD;
    const bf16* gk = gq + D;
    const bf16* gv = gq + 2 * D;
    const bf16* gdo = d_out + static_cast<size_t>(n) * T * D + h * HD;
    const float* glse = lse + (static_cast<size_t>(n) * heads + h) * T;
    const float* gdel = delta + (static_cast<size_t>(n) * heads + h) * T;
    const int nkv = (T + kTile - 1) / kTile;

#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
        load_tile_async(sm.kv[0][blk], gq + blk * kHd, ld, q0, T);
        load_tile_async(sm.kv[1][blk], gdo + blk * kHd, D, q0, T);
        load_tile_async(sm.q[0][blk], gk + blk * kHd, ld, 0, T);
        load_tile_async(sm.d_o[0][blk], gv + blk * kHd, ld, 0, T);
    }
    cp_async_commit();

    const bool warp_active = (q0 + warp * 16) < T;
    const int r0 = q0 + warp * 16 + (lane >> 2);
    float lse_r[2], del_r[2];
    lse_r[0] = (r0 < T) ? glse[r0] : INFINITY;
    lse_r[1] = (r0 + 8 < T) ? glse[r0 + 8] : INFINITY;
    del_r[0] = (r0 < T) ? gdel[r0] : 0.f;
    del_r[1] = (r0 + 8 < T) ? gdel[r0 + 8] : 0.f;

    uint32_t qf[NB][4][4], dof[NB][4][4];
    float dq[NB * 8][4];
#pragma unroll
    for (int i = 0; i < NB * 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;

    for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j + 1 < nkv) {
#pragma unroll
            for (int blk = 0; blk < NB; ++blk) {
                load_tile_async(sm.q[buf ^ 1][blk], gk + blk * kHd, ld, (j + 1) * kTile, T);
                load_tile_async(sm.d_o[buf ^ 1][blk], gv + blk * kHd, ld, (j + 1) * kTile, T);
            }
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            if (j == 0) {
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) {
                    load_a_frags(qf[blk], sm.kv[0][blk], warp * 16);
                    load_a_frags(dof[blk], sm.kv[1][blk], warp * 16);
                }
            }
            const int valid = min(kTile, T - j * kTile);
            const int npair = (valid + 15) >> 4;
            float s[8][4], dp[8][4];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) s[a][c] = dp[a][c] = 0.f;
#pragma unroll
            for (int pair = 0; pair < 4; ++pair) {
                if (pair < npair) {
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            uint32_t b[4];
                            load_b_nk(b, sm.q[buf][blk], pair, ks);  // K tile
                            mma_bf16_16816(s[2 * pair], qf[blk][ks], b[0], b[1]);
                            mma_bf16_16816(s[2 * pair + 1], qf[blk][ks], b[2], b[3]);
                            load_b_nk(b, sm.d_o[buf][blk], pair, ks);  // V tile
                            mma_bf16_16816(dp[2 * pair], dof[blk][ks], b[0], b[1]);
                            mma_bf16_16816(dp[2 * pair + 1], dof[blk][ks], b[2], b[3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int col = nt * 8 + ((lane & 3) << 1);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool ok = (col + (c & 1)) < valid;
                    const float pv = ok ? exp2f((s[nt][c] - lse_r[c >> 1]) * kLog2e) : 0.f;
                    s[nt][c] = pv * (dp[nt][c] - del_r[c >> 1]);  // dS
                }
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (ks < npair) {
                    uint32_t da[4];
                    da[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
                    da[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
                    da[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
                    da[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
#pragma unroll
                        for (int dpi = 0; dpi < 4; ++dpi) {
                            uint32_t b[4];
                            load_b_kn(b, sm.q[buf][blk], dpi, ks);  // K tile as [key, d]
                            mma_bf16_16816(dq[blk * 8 + 2 * dpi], da, b[0], b[1]);
                            mma_bf16_16816(dq[blk * 8 + 2 * dpi + 1], da, b[2], b[3]);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (warp_active) {
        bf16* gdq = d_qkv + static_cast<size_t>(n) * T * ld + h * HD;
#pragma unroll
        for (int blk = 0; blk < NB; ++blk)
            store_tile_bf16(reinterpret_cast<const float(&)[8][4]>(dq[blk * 8]), sm.kv[0][blk], warp * 16, gdq + blk * kHd, ld,
                            q0, T);
    }
}

template <int NB>
int fwd_legacy_nb(const void* qkv, void* out, float* lse, int n, int T, int heads, int q_begin, cudaStream_t s) {
    constexpr int kSmem = 5 * NB * kTile * kHd * 2;  // Q + 2 x (K, V) stages
    static PerDeviceOnce configured;
    PCG_ONCE_PER_DEVICE(configured, PCG_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)));
    const dim3 grid(ceil_div(T - q_begin, kTile), heads, n);
    attn_fwd_kernel<NB><<<grid, 128, kSmem, s>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, T, heads, q_begin);
    PCG_LAUNCH_CHECK("attn_fwd_kernel");
    return 0;
}

template <int NB>
int delta_nb(const void* out, const void* d_out, float* delta, int n, int T, int heads, cudaStream_t s) {
    const int rows = n * T;
    const long long threads = static_cast<long long>(rows) * heads * 8 * NB;
    attn_delta_kernel<NB><<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(
        static_cast<const bf16*>(out), static_cast<const bf16*>(d_out), delta, rows, T, heads);
    PCG_LAUNCH_CHECK("attn_delta_kernel");
    return 0;
}

template <int NB>
int bwd_legacy_nb(const void* qkv, const void* d_out, const float* lse, const float* delta, void* d_qkv, int n, int T,
                  int heads, int begin, cudaStream_t s) {
    constexpr int kSmem = static_cast<int>(sizeof(BwdSmem<NB>));
    static PerDeviceOnce configured;
    PCG_ONCE_PER_DEVICE(configured, PCG_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); PCG_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)));
    const dim3 grid(ceil_div(T - begin, kTile), heads, n);
    attn_bwd_dkdv_kernel<NB><<<grid, 128, kSmem, s>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(d_out), lse,
                                                     delta, static_cast<bf16*>(d_qkv), T, heads, begin);
    PCG_LAUNCH_CHECK("attn_bwd_dkdv_kernel");
    attn_bwd_dq_kernel<NB><<<grid, 128, kSmem, s>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(d_out), lse, delta,
                                                   static_cast<bf16*>(d_qkv), T, heads, begin);
    PCG_LAUNCH_CHECK("attn_bwd_dq_kernel");
    return 0;
}

}  // namespace
}  // namespace pcg

namespace pcg {

// mma.sync path for query rows [q_begin, T) (everything when q_begin == 0)
int attn_fwd_legacy(const void* qkv, void* out, float* lse, int n, int T, int heads, int q_begin, cudaStream_t s) {
    if (q_begin >= T) return 0;
    return fwd_legacy_nb<1>(qkv, out, lse, n, T, heads, q_begin, s);
}

int attn_delta(const void* out, const void* d_out, float* delta, int n, int T, int heads, cudaStream_t s) {
    return delta_nb<1>(out, d_out, delta, n, T, heads, s);
}

// mma.sync backward for key rows [k_begin, T) (dK, dV) and query rows [q_begin, T) (dQ); delta must be ready
int attn_bwd_legacy(const void* qkv, const void* d_out, const float* lse, const float* delta, void* d_qkv, int n, int T,
                    int heads, int begin, cudaStream_t s) {
    if (begin >= T) return 0;
    return bwd_legacy_nb<1>(qkv, d_out, lse, delta, d_qkv, n, T, heads, begin, s);
}

}  // namespace pcg

using namespace pcg;

// Wide heads (head dim 80 / 88 padded to 128 columns per head: ViT-H/14, ViT-g/14, the reference's OpenCLIP default,
// perceptor/losses/open_clip.py:8-12): qkv is [n*T, 3 * heads * 128], out / d_out [n*T, heads * 128]; the pad columns
// are zero on input (zero weight rows) and come out zero.  mma.sync kernels: these towers are outside BASELINE.json's
// configs, so they get a correct path, not a tcgen05 one.
extern "C" int pcg_attn_fwd_wide(const void* qkv, void* out, float* lse, int n, int T, int heads, void* stream) {
    PCG_CHECK_ARG(qkv && out && lse, "pcg_attn_fwd_wide: null pointer");
    PCG_CHECK_ARG(n > 0 && T > 0 && heads > 0 && n <= 65535 && heads <= 65535, "pcg_attn_fwd_wide: bad shape n=%d T=%d heads=%d", n, T, heads);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_FWD, 4.0 * T * T * 128 * heads * n, s);
    return fwd_legacy_nb<2>(qkv, out, lse, n, T, heads, 0, s);
}

extern "C" int pcg_attn_bwd_wide(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                                 void* d_qkv, int n, int T, int heads, void* stream) {
    PCG_CHECK_ARG(qkv && out && d_out && lse && delta_ws && d_qkv, "pcg_attn_bwd_wide: null pointer");
    PCG_CHECK_ARG(n > 0 && T > 0 && heads > 0 && n <= 65535 && heads <= 65535, "pcg_attn_bwd_wide: bad shape n=%d T=%d heads=%d", n, T, heads);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfileScope prof(PCG_PROF_ATTN_BWD, 8.0 * T * T * 128 * heads * n, s);
    if (int rc = delta_nb<2>(out, d_out, delta_ws, n, T, heads, s)) return rc;
    return bwd_legacy_nb<2>(qkv, d_out, lse, delta_ws, d_qkv, n, T, heads, 0, s);
}
