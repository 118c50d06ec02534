// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  out[M,N] = epilogue( A[M,K] * B[N,K]^T ).
//
//   * A and B are bf16, K-major (row-major [rows, K]) — exactly nn.Linear's activation / weight layout, so the
//     forward GEMMs consume weights as stored and the dgrad GEMMs consume pre-transposed copies.
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) stages 128 x 64 (A) and BN x 64 (B) tiles through an
//     mbarrier ring; one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered
//     fp32 accumulator in tensor memory; eight epilogue warps drain it with tcgen05.ld, transpose each 32 x 32
//     chunk through a swizzled shared-memory tile and apply the fused epilogue (bias / QuickGELU / residual add /
//     activation derivative) with fully coalesced global loads and stores; epilogue inputs are prefetched ahead.
//   * CTAS = 2: a cluster of two CTAs (one TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2: each CTA
//     stages its own 128 rows of A and HALF of the B tile, so shared memory sees 2/3 of the single-CTA traffic
//     per flop (the single-CTA 128 x 256 tile is shared-memory-bandwidth bound at ~75 % of the MMA rate, see
//     tools/ubench.cu); the even CTA issues the MMAs, each CTA drains its own 128 accumulator rows.
//   * grid = min(#tiles, #SMs); each CTA walks tiles t = blockIdx.x + i*gridDim.x (n fastest so that the CTAs
//     running concurrently share A row-panels in L2; B (weights) is L2 resident).
//
// Replaces: nn.Linear forward/dgrad and conv1-as-GEMM in the reference ViT (perceptor/models/ruclip/model.py:
// 31-39, 43-49, 85-91), which today are eager cuBLAS calls plus separate bias/activation/residual kernels.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "pcg_common.cuh"
#include "pcg_ptx.cuh"

namespace pcg {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;   // warps 0-7: two warpgroups; warp % 4 selects the TMEM lane quarter, warp / 4 the column half
constexpr int kCtrlWarps = 3;  // warp 8: TMA producer, warp 9: MMA issuer, warp 10: TMEM allocator
constexpr int kWarpTma = kEpiWarps, kWarpMma = kEpiWarps + 1, kWarpAlloc = kEpiWarps + 2;
constexpr int kThreads = 32 * (kCtrlWarps + kEpiWarps);  // 352: leaves 184 registers per thread for the epilogue
constexpr int kABytes = BM * BK * 2;
// A stays evict-normal: every row panel of A is read by all N / 256 column tiles, and marking it evict-first (to keep the
// output rows in L2 for the LayerNorm that follows) cost 0.9 ms per ViT-L/14 step.
constexpr uint64_t kHintA = kEvictNormal;
constexpr int kEpiStageBytes = 32 * 32 * 4;  // per epilogue warp: one 32 x 32 fp32 chunk, XOR-swizzled 16-byte columns

template <int BN, int CTAS = 1>
struct TileCfg {
    static constexpr int kBBytes = BN / CTAS * BK * 2;  // per CTA: a pair splits the B tile by rows
    static constexpr int kStages = (CTAS == 2) ? 6 : (BN == 256) ? 4 : (BN == 192) ? 4 : (BN == 128) ? 6 : 8;
    static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int kBarBytes = 256;
    static constexpr int kEpiBytes = kEpiWarps * kEpiStageBytes;
    static constexpr int kSmem = kStages * (kABytes + kBBytes) + kEpiBytes + kBarBytes + 1024;  // +1024: manual alignment
    static_assert(kSmem <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

struct GemmParams {
    int M, N, K;
    const float* bias;
    const void* aux;
    void* out;
    void* out2;
    int ldo;
    int act;
};

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// QuickGELU x * sigmoid(1.702 x) with sigmoid(z) = 0.5 * tanh(z / 2) + 0.5: one MUFU.TANH + 3 FMA-class ops.
// The activation is a template parameter so that erff never bloats the QuickGELU kernels' instruction footprint.
// Returns act(h) and writes act'(h): the forward epilogue stores the DERIVATIVE (not the pre-activation) for the
// backward pass, which then only multiplies (dgrad-only backward never needs h itself).
template <int ACT>
__device__ __forceinline__ float act_and_grad(float h, float& grad) {
    if constexpr (ACT == PCG_ACT_QUICKGELU) {
        const float s = fmaf(0.5f, tanh_approx(0.851f * h), 0.5f);  // sigmoid(1.702 h)
        grad = s * fmaf(1.702f * h, 1.0f - s, 1.0f);
        return h * s;
    } else {
        const float cdf = 0.5f * (1.0f + erff(h * 0.70710678118654752f));
        grad = cdf + h * 0.3989422804014327f * __expf(-0.5f * h * h);
        return h * cdf;
    }
}

// Two columns at a time on the packed fp32 pipe (FFMA2 / FMUL2 / FADD2: one issue slot for two lanes of fp32 math).
// The fc epilogue is bound by the FMA pipe's issue rate next to the MMA main loop (11 fp32 instructions per element
// in scalar form against a K = 1024 main loop); in packed form and with the derivative written as
//   act'(h) = s + 0.4255 h (1 - t^2),  t = tanh(0.851 h),  s = 0.5 + 0.5 t   (1.702 s (1 - s) = 0.4255 (1 - t^2))
// it is 3.5 FMA-pipe instructions + one MUFU per element.
template <int ACT>
__device__ __forceinline__ float2 act_and_grad2(float2 h, float2& grad) {
    if constexpr (ACT == PCG_ACT_QUICKGELU) {
        const float2 z = __fmul2_rn(h, make_float2(0.851f, 0.851f));
        const float2 t = make_float2(tanh_approx(z.x), tanh_approx(z.y));
        const float2 s = __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
        const float2 u = __ffma2_rn(t, t, make_float2(-1.0f, -1.0f));  // t^2 - 1: the sign goes into the constant below
        grad = __ffma2_rn(__fmul2_rn(h, u), make_float2(-0.4255f, -0.4255f), s);
        return __fmul2_rn(h, s);
    } else {
        float2 a;
        a.x = act_and_grad<ACT>(h.x, grad.x);
        a.y = act_and_grad<ACT>(h.y, grad.y);
        return a;
    }
}

// MC (CTA pairs only): clusters of FOUR CTAs = two pairs that work on the same 256 output rows and on adjacent column
// tiles (n, n + 1).  Both pairs need the same A rows, so every CTA fetches 64 of its pair-half's 128 rows and
// multicasts them to the CTA of the same rank in the other pair: a k-block costs an SM 8 KB of A + 16 KB of B from L2
// instead of 16 + 16.  (The GEMMs of the step sit at the L2's practical throughput: their rates follow
// operand + epilogue bytes per tile, see DESIGN.md.)  A stage may be refilled once BOTH pairs' MMAs have read it: the
// empty barriers collect one tcgen05.commit from each pair leader.
// PCG_GEMM_TS (compile time): the bf16-output epilogues (BF16, BIAS_ACT) keep the tensor-memory layout (thread = row),
// write packed bf16 rows into two 2 KB staging tiles per warp (64-byte swizzle) and send them out with bulk tensor
// stores, instead of transposing fp32 through shared memory for coalesced st.global: per 32-column chunk and warp that
// is 16-32 L1 wavefronts instead of 96-128.
#ifndef PCG_GEMM_TS
#define PCG_GEMM_TS 1
#endif

template <int BN, int MODE, int ACT, int CTAS, bool MC = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_o2,
                    const GemmParams p) {
    static_assert(!MC || CTAS == 2, "multicast clusters are built from CTA pairs");
    using Cfg = TileCfg<BN, CTAS>;
    constexpr int kCluster = CTAS * (MC ? 2 : 1);
    constexpr int kTileM = BM * CTAS;  // rows of one output tile (per pair / CTA)
    constexpr int kStages = Cfg::kStages;
    constexpr int kBBytes = Cfg::kBBytes;

    extern __shared__ uint8_t smem_raw[];
    // 128-byte swizzle atoms need 1024-byte aligned stage buffers.
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kABytes;
    uint8_t* smem_epi = smem + kStages * (kABytes + kBBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tfull_bar = bars + 2 * kStages;
    uint64_t* tempty_bar = bars + 2 * kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (p.M + kTileM - 1) / kTileM;
    const int num_n = (p.N + BN - 1) / BN;
    const int num_kb = (p.K + BK - 1) / BK;
    // a CTA pair walks the same tiles; rank 1 owns rows [128, 256) of the tile and B rows [BN/2, BN)
    const int crank = (CTAS == 2) ? static_cast<int>(cluster_ctarank()) : 0;
    const int cta_rank = crank & 1;
    const int pair = MC ? (crank >> 1) : 0;
    // work units of a cluster: output tiles, or (MC) a row tile x two adjacent column tiles, one per pair
    const int units_n = MC ? (num_n >> 1) : num_n;
    const int num_tiles = num_m * units_n;
    const int tile0 = static_cast<int>(blockIdx.x) / kCluster;
    const int tile_step = static_cast<int>(gridDim.x) / kCluster;
    auto tile_m = [&](int t) { return t / units_n; };
    auto tile_n = [&](int t) { return MC ? 2 * (t % units_n) + pair : t % units_n; };

    if (warp == kWarpTma && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == kWarpMma && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], MC ? 2 : 1);  // MC: one commit from each pair of the cluster
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], kEpiWarps * CTAS);  // the even CTA's copy collects both CTAs' epilogue warps
        }
        fence_barrier_init();
    }
    if (warp == kWarpAlloc) {
        if constexpr (CTAS == 2) {
            tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_slot, Cfg::kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();  // barriers visible to the peer before use
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the previous kernel's tail;
    // nothing below may read what it produced before it has completed
    grid_dep_wait();
    grid_dep_launch();

    if (warp == kWarpTma) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                const int m0 = tile_m(tile) * kTileM + cta_rank * BM;
                const int n0 = tile_n(tile) * BN + cta_rank * (BN / CTAS);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (MC) {
                        // rows [64 pair, 64 pair + 64) of this CTA's 128 A rows, to this CTA and to its twin in the other pair
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + kBBytes));
                        tma_load_2d_2sm_mc(&map_a, &full_bar[stage], smem_a + stage * kABytes + pair * (kABytes / 2), kb * BK,
                                           m0 + pair * (BM / 2), static_cast<uint16_t>(0x5u << cta_rank), kHintA);
                        tma_load_2d_2sm(&map_b, &full_bar[stage], smem_b + stage * kBBytes, kb * BK, n0, kEvictLast);
                    } else if constexpr (CTAS == 2) {
                        // both CTAs' bytes land on the even CTA's barrier, which alone expects them
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + kBBytes));
                        tma_load_2d_2sm(&map_a, &full_bar[stage], smem_a + stage * kABytes, kb * BK, m0, kHintA);
                        tma_load_2d_2sm(&map_b, &full_bar[stage], smem_b + stage * kBBytes, kb * BK, n0, kEvictLast);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], kABytes + kBBytes);
                        tma_load_2d(&map_a, &full_bar[stage], smem_a + stage * kABytes, kb * BK, m0, kHintA);
                        tma_load_2d(&map_b, &full_bar[stage], smem_b + stage * kBBytes, kb * BK, n0, kEvictLast);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t desc_a = umma_smem_desc_sw128(smem_u32(smem_a + stage * kABytes));
                    const uint64_t desc_b = umma_smem_desc_sw128(smem_u32(smem_b + stage * kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 bytes per K=16 step inside the 128-byte swizzle atom (encoded >>4 => +2)
                        if constexpr (CTAS == 2)
                            umma_f16_2sm(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0);
                        else
                            umma_f16(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0);
                    }
                    if constexpr (CTAS == 2) {
                        // frees the slot in both CTAs (MC: in all four) when these MMAs retire
                        umma_commit_2sm_mask(&empty_bar[stage], MC ? 0xF : 0x3);
                        if (kb == num_kb - 1) umma_commit_2sm_mask(&tfull_bar[as], static_cast<uint16_t>(0x3u << (2 * pair)));
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp < kEpiWarps && PCG_GEMM_TS && (MODE == PCG_GEMM_BF16 || MODE == PCG_GEMM_BIAS_ACT)) {
        // ===================== epilogue with bulk tensor stores (thread = row throughout) =====================
        const int quarter = warp & 3, half = warp >> 2;
        constexpr int kHalfN = BN / 2, kChunks = kHalfN / 32;
        uint8_t* buf0 = smem_epi + warp * kEpiStageBytes;  // [32 rows x 64 B], 16-byte chunk k of row r at k ^ ((r >> 1) & 3)
        uint8_t* buf1 = buf0 + 2048;
        const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
        [[maybe_unused]] uint32_t n_put = 0;
        int it = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int row_base = tile_m(tile) * kTileM + cta_rank * BM + quarter * 32;
            const int col_base = tile_n(tile) * BN + half * kHalfN;
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * kHalfN;
            uint32_t r[32];
            tmem_ld_32x32(taddr0, r);
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int col0 = col_base + c * 32;
                float4 b4[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)  // the same 16 bytes for every lane: one wavefront, an L1 hit after the first warp
                    b4[j] = (p.bias != nullptr && col0 < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                tmem_wait_ld();
                uint32_t pk[16], pg[MODE == PCG_GEMM_BIAS_ACT ? 16 : 1];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 v01 = __fadd2_rn(make_float2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])),
                                                  make_float2(b4[j].x, b4[j].y));
                    const float2 v23 = __fadd2_rn(make_float2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])),
                                                  make_float2(b4[j].z, b4[j].w));
                    if constexpr (MODE == PCG_GEMM_BIAS_ACT) {
                        float2 g01, g23;
                        const float2 a01 = act_and_grad2<ACT>(v01, g01);
                        const float2 a23 = act_and_grad2<ACT>(v23, g23);
                        pk[2 * j] = pack_bf16(a01.x, a01.y), pk[2 * j + 1] = pack_bf16(a23.x, a23.y);
                        pg[2 * j] = pack_bf16(g01.x, g01.y), pg[2 * j + 1] = pack_bf16(g23.x, g23.y);
                    } else {
                        pk[2 * j] = pack_bf16(v01.x, v01.y), pk[2 * j + 1] = pack_bf16(v23.x, v23.y);
                    }
                }
                if (c + 1 < kChunks) {
                    tmem_ld_32x32(taddr0 + (c + 1) * 32, r);  // in flight while this chunk goes out
                } else {
                    // the accumulator is in registers: the MMA thread may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CTAS == 2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]);
                    }
                }
                // Two staging tiles, every store its own bulk group.  A tile may be rewritten once the store issued from it
                // two groups ago has read it -- "at most one group pending": BF16 alternates between the tiles from chunk
                // to chunk, BIAS_ACT sends act'(h) through tile 0 and act(h) through tile 1 in every chunk.
                auto put = [&](uint8_t* dst, const uint32_t* q, const CUtensorMap* map) {
                    if (lane == 0) bulk_wait_read<1>();
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<uint4*>(dst + static_cast<uint32_t>(lane) * 64u + ((static_cast<uint32_t>(k) ^ sw) << 4)) =
                            make_uint4(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && col0 < p.N) {
                        tma_store_2d(map, dst, col0, row_base);
                        bulk_commit_group();
                    }
                };
                if constexpr (MODE == PCG_GEMM_BIAS_ACT) {
                    put(buf0, pg, &map_o);   // out  = act'(h)
                    put(buf1, pk, &map_o2);  // out2 = act(h)
                } else {
                    // (a running count, not the chunk index: with an odd number of chunks per tile -- BN = 64, 192 -- the
                    // first chunk of the next tile would otherwise reuse the tile that the store just before it reads)
                    put((n_put++ & 1) == 0 ? buf0 : buf1, pk, &map_o);
                }
            }
        }
        if (lane == 0) bulk_wait_all();
    } else if (warp < kEpiWarps) {
        // TMEM -> registers (thread = row) -> XOR-swizzled per-warp staging tile in smem -> registers
        // (8 lanes = one row's 32 columns) so that every global access is a full, coalesced row segment.
        // Epilogue inputs (residual / pre-activation) are prefetched kDist 32-column chunks ahead, across tile
        // boundaries, to keep >= 64 KB of reads in flight per SM (the epilogue is latency-, not throughput-bound).
        const int quarter = warp & 3;               // TMEM lanes [32*quarter, +32) are the ones this warp may read
        const int half = warp >> 2;                 // column half of the tile
        constexpr int kHalfN = BN / 2;
        constexpr int kChunks = kHalfN / 32;
        constexpr bool kHasAux = (MODE == PCG_GEMM_RESID_F32 || MODE == PCG_GEMM_DACT);
        constexpr int kAuxWords = (MODE == PCG_GEMM_RESID_F32) ? 4 : 2;  // 32-bit words per lane per row group
        constexpr int kDistWanted = 2;  // deeper prefetch spills: 2 slots x 8 rows x 16 B (or 8 B) per lane fit the register file
        // the slot of chunk c must be a compile-time constant: kDist has to divide kChunks
        constexpr int kDist = !kHasAux ? 1 : (kChunks % kDistWanted == 0 ? kDistWanted : (kChunks % 2 == 0 ? 2 : kChunks));
        uint8_t* stage_buf = smem_epi + warp * kEpiStageBytes;
        const int sub_row = lane >> 3;  // row within a 4-row group when reading back
        const int sub_col = lane & 7;   // 16-byte column chunk (4 floats) within the 32-column chunk
        uint32_t auxr[kDist][8][kAuxWords];
        // issue the loads of chunk `c` of tile `t` into slot `c % kDist`
        // Loads are unconditional (indices clamped into the matrix, the consumer is guarded): a predicated load would
        // land in a temporary and the copy into the slot would wait for it right here, defeating the prefetch.
        auto prefetch_aux = [&](int t, int c, uint32_t(&dst)[8][kAuxWords]) {
            if constexpr (kHasAux) {
                const int tt = min(t, num_tiles - 1);
                const int rb = tile_m(tt) * kTileM + cta_rank * BM + quarter * 32 + sub_row;
                const int col = min(tile_n(tt) * BN + half * kHalfN + c * 32 + sub_col * 4, p.N - 4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const size_t off = static_cast<size_t>(min(rb + i * 4, p.M - 1)) * p.ldo + col;
                    if constexpr (MODE == PCG_GEMM_RESID_F32) {
                        const uint4 t4 = *reinterpret_cast<const uint4*>(static_cast<const float*>(p.aux) + off);
                        dst[i][0] = t4.x; dst[i][1] = t4.y; dst[i][kAuxWords - 2] = t4.z; dst[i][kAuxWords - 1] = t4.w;
                    } else {
                        const uint2 t2 = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p.aux) + off);
                        dst[i][0] = t2.x; dst[i][1] = t2.y;
                    }
                }
            }
        };
#pragma unroll
        for (int d = 0; d < kDist; ++d) prefetch_aux(tile0 + (d / kChunks) * tile_step, d % kChunks, auxr[d % kDist]);
        int it = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = tile_m(tile) * kTileM + cta_rank * BM;
            const int n0 = tile_n(tile) * BN;
            const int row_base = m0 + quarter * 32;
            const int col_base = n0 + half * kHalfN + sub_col * 4;  // this lane's first column of chunk 0
            // Bias: requested one 32-column chunk ahead (chunk 0 before the wait for the accumulator).  A load issued in
            // the chunk that uses it sits exposed -- an L2 round trip, four times per tile -- on the chain that has to stay
            // shorter than the next tile's main loop.  DACT (a dgrad epilogue) never has a bias.
            constexpr bool kMayBias = MODE != PCG_GEMM_DACT;
            auto load_bias = [&](int c) {
                const int col = col_base + c * 32;
                return (kMayBias && p.bias != nullptr && col < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col))
                                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            float4 bias_cur = load_bias(0);
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * kHalfN;
            uint32_t r[32];
            tmem_ld_32x32(taddr0, r);
#pragma unroll  // fully unrolled: registers of an in-flight tcgen05.ld must not be carried around a loop back-edge
            for (int c = 0; c < kChunks; ++c) {
                const int col = col_base + c * 32;
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(stage_buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                        make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                if (c + 1 < kChunks) tmem_ld_32x32(taddr0 + (c + 1) * 32, r);  // in flight while chunk c is written out
                __syncwarp();
                const bool col_ok = col < p.N;
                const float4 bias4 = bias_cur;
                if (c + 1 < kChunks) bias_cur = load_bias(c + 1);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lrow = i * 4 + sub_row;
                    const int grow = row_base + lrow;
                    float4 v = *reinterpret_cast<const float4*>(stage_buf + lrow * 128 + ((sub_col ^ (lrow & 7)) << 4));
                    if (grow < p.M && col_ok) {
                        {
                            const float2 v01 = __fadd2_rn(make_float2(v.x, v.y), make_float2(bias4.x, bias4.y));
                            const float2 v23 = __fadd2_rn(make_float2(v.z, v.w), make_float2(bias4.z, bias4.w));
                            v = make_float4(v01.x, v01.y, v23.x, v23.y);
                        }
                        const size_t off = static_cast<size_t>(grow) * p.ldo + col;
                        const uint32_t(&ax)[kAuxWords] = auxr[c % kDist][i];
                        if constexpr (MODE == PCG_GEMM_BF16) {
                            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + off) =
                                make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
                        } else if constexpr (MODE == PCG_GEMM_BIAS_ACT) {
                            float2 g01, g23;
                            const float2 a01 = act_and_grad2<ACT>(make_float2(v.x, v.y), g01);
                            const float2 a23 = act_and_grad2<ACT>(make_float2(v.z, v.w), g23);
                            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + off) =
                                make_uint2(pack_bf16(g01.x, g01.y), pack_bf16(g23.x, g23.y));
                            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out2) + off) =
                                make_uint2(pack_bf16(a01.x, a01.y), pack_bf16(a23.x, a23.y));
                        } else if constexpr (MODE == PCG_GEMM_RESID_F32) {
                            const float2 r01 = __fadd2_rn(make_float2(__uint_as_float(ax[0]), __uint_as_float(ax[1])),
                                                          make_float2(v.x, v.y));
                            const float2 r23 = __fadd2_rn(make_float2(__uint_as_float(ax[kAuxWords - 2]),
                                                                      __uint_as_float(ax[kAuxWords - 1])), make_float2(v.z, v.w));
                            *reinterpret_cast<float4*>(static_cast<float*>(p.out) + off) = make_float4(r01.x, r01.y, r23.x, r23.y);
                        } else if constexpr (MODE == PCG_GEMM_DACT) {
                            const float2 d01 = __fmul2_rn(make_float2(v.x, v.y), make_float2(__uint_as_float(ax[0] << 16),
                                                                                            __uint_as_float(ax[0] & 0xffff0000u)));
                            const float2 d23 = __fmul2_rn(make_float2(v.z, v.w), make_float2(__uint_as_float(ax[1] << 16),
                                                                                            __uint_as_float(ax[1] & 0xffff0000u)));
                            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + off) =
                                make_uint2(pack_bf16(d01.x, d01.y), pack_bf16(d23.x, d23.y));
                        } else {  // PCG_GEMM_F32
                            *reinterpret_cast<float4*>(static_cast<float*>(p.out) + off) = v;
                        }
                    }
                }
                // slot c % kDist is free again: prefetch the chunk kDist items ahead (possibly of a later tile)
                prefetch_aux(tile + ((c + kDist) / kChunks) * tile_step, (c + kDist) % kChunks, auxr[c % kDist]);
                __syncwarp();  // staging tile is rewritten by the next chunk
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTAS == 2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]);
            }
        }
    }

    tc_fence_before();
    if constexpr (CTAS == 2) {
        cluster_sync_all();  // neither CTA may exit while the other can still signal its barriers
        if (warp == kWarpAlloc) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    } else {
        __syncthreads();
        if (warp == kWarpAlloc) tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------
using EncodeFn = PFN_cuTensorMapEncodeTiled_v12000;

EncodeFn get_encode_fn() {
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

struct MapKey {
    const void* ptr;
    int rows, cols, ld, box_rows;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ static_cast<size_t>(k.rows);
        h = h * 1000003u ^ static_cast<size_t>(k.cols);
        h = h * 1000003u ^ static_cast<size_t>(k.ld);
        h = h * 1000003u ^ static_cast<size_t>(k.box_rows);
        return h;
    }
};

// store map of a bf16 [rows, cols] output (row stride ld): 32 x 32 boxes, 64-byte swizzle (PCG_GEMM_TS epilogue)
int get_store_map(CUtensorMap* out, const void* ptr, int rows, int cols, int ld) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, rows, cols, ld, -32};
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeFn encode = get_encode_fn();
    if (encode == nullptr) return set_error(-2, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(-3, "cuTensorMapEncodeTiled(store) failed (CUresult %d) rows=%d cols=%d ld=%d ptr=%p",
                         static_cast<int>(r), rows, cols, ld, ptr);
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

int get_tensor_map(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int box_rows) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeFn encode = get_encode_fn();
    if (encode == nullptr) return set_error(-2, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    const cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(-3, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%d cols=%d ld=%d box_rows=%d ptr=%p",
                         static_cast<int>(r), rows, cols, ld, box_rows, ptr);
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

// Programmatic dependent launch of the GEMMs (their prologue runs under the previous kernel's tail): on by default,
// PCG_PDL=0 turns it off.  Same-box A/B: 0.3-0.6 % of a 128-cutout step, ~1 % of a 16-cutout step.
bool g_pdl_enabled = []() {
    const char* e = getenv("PCG_PDL");
    return !(e != nullptr && e[0] == '0');
}();

// clusters of four that the device keeps resident at once (GPCs with 18 SMs hold four of them and leave a TPC idle)
template <typename Kernel>
int max_clusters_of_4(Kernel kernel, int smem, int* out) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    PCG_CUDA(cudaOccupancyMaxActiveClusters(&n, kernel, &cfg));
    *out = n;
    return 0;
}

template <int BN, int MODE, int ACT, int CTAS, bool MC = false>
int launch_gemm_a(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& mo2,
                  const GemmParams& p, cudaStream_t stream) {
    using Cfg = TileCfg<BN, CTAS>;
    constexpr int kCluster = CTAS * (MC ? 2 : 1);
    static PerDeviceOnce configured;
    PCG_ONCE_PER_DEVICE(configured, PCG_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, MODE, ACT, CTAS, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem)));
    const int tiles = ceil_div(p.M, BM * CTAS) * (MC ? ceil_div(p.N, BN) / 2 : ceil_div(p.N, BN));
    int slots = sm_count() / kCluster;  // clusters (or CTAs) that run concurrently
    if constexpr (MC) {
        static int resident[64] = {0};
        int dev = 0;
        PCG_CUDA(cudaGetDevice(&dev));
        if (dev < 64 && resident[dev] == 0) {
            int n = 0;
            if (int rc = max_clusters_of_4(gemm_tcgen05_kernel<BN, MODE, ACT, CTAS, MC>, Cfg::kSmem, &n)) return rc;
            resident[dev] = n > 0 ? n : 1;
        }
        if (dev < 64) slots = resident[dev] < slots ? resident[dev] : slots;
    }
    const int grid = (tiles < slots ? tiles : slots) * kCluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if constexpr (CTAS == 2) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = kCluster;
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    if (g_pdl_enabled) {  // start this grid's prologue under the previous kernel's tail (see grid_dep_wait)
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    PCG_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, MODE, ACT, CTAS, MC>, ma, mb, mo, mo2, p));
    PCG_LAUNCH_CHECK("gemm_tcgen05_kernel");
    return 0;
}
template <int BN, int MODE, int CTAS, bool MC = false>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& mo2,
                const GemmParams& p, cudaStream_t stream) {
    if constexpr (MODE == PCG_GEMM_BIAS_ACT) {
        if (p.act == PCG_ACT_GELU) return launch_gemm_a<BN, MODE, PCG_ACT_GELU, CTAS, MC>(ma, mb, mo, mo2, p, stream);
    }
    return launch_gemm_a<BN, MODE, PCG_ACT_QUICKGELU, CTAS, MC>(ma, mb, mo, mo2, p, stream);
}

template <int BN, int CTAS = 1, bool MC = false>
int dispatch_mode(int mode, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& mo2,
                  const GemmParams& p, cudaStream_t s) {
    switch (mode) {
        case PCG_GEMM_BF16: return launch_gemm<BN, PCG_GEMM_BF16, CTAS, MC>(ma, mb, mo, mo2, p, s);
        case PCG_GEMM_BIAS_ACT: return launch_gemm<BN, PCG_GEMM_BIAS_ACT, CTAS, MC>(ma, mb, mo, mo2, p, s);
        case PCG_GEMM_RESID_F32: return launch_gemm<BN, PCG_GEMM_RESID_F32, CTAS, MC>(ma, mb, mo, mo2, p, s);
        case PCG_GEMM_DACT: return launch_gemm<BN, PCG_GEMM_DACT, CTAS, MC>(ma, mb, mo, mo2, p, s);
        case PCG_GEMM_F32: return launch_gemm<BN, PCG_GEMM_F32, CTAS, MC>(ma, mb, mo, mo2, p, s);
        default: return set_error(-1, "pcg_gemm_bf16: unknown mode %d", mode);
    }
}

// Pick the N tile that minimises (waves x per-tile cost).  Per-tile MMA time is proportional to BN; narrow
// tiles pay more shared-memory traffic per flop, hence the mild penalty.
// Returns the tile width; *cost_out = modelled time in units of (one SM computing 128 rows x 1 column).
int choose_bn(int M, int N, float* cost_out = nullptr) {
    const int sms = sm_count();
    const int cands[4] = {256, 192, 128, 64};
    const float penalty[4] = {1.00f, 1.03f, 1.10f, 1.45f};
    int best = 128;
    float best_cost = 1e30f;
    for (int i = 0; i < 4; ++i) {
        const int bn = cands[i];
        const int tiles = ceil_div(M, BM) * ceil_div(N, bn);
        const int waves = ceil_div(tiles, sms);
        // wasted columns in the last n tile still cost MMA time
        const float cost = static_cast<float>(waves) * bn * penalty[i];
        if (cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    if (cost_out != nullptr) *cost_out = best_cost;
    return best;
}
// The same model for CTA pairs (256 x 256 tiles over sms / 2 clusters): each SM of a pair computes 128 x 256 per
// tile at ~0.92 of the single-CTA cost (measured in-step: 1210 vs 1113 TFLOP/s), but the wave count is quantised
// on half as many slots -- with few tiles (ViT-B/32: 150 tiles of N = 768 on 74 pairs = 2.03 waves) a narrower
// single-CTA tile wins.
float pair_cost(int M, int N) {
    const int waves = ceil_div(ceil_div(M, 256) * (N / 256), sm_count() / 2);
    return static_cast<float>(waves) * 256.0f * 0.92f;
}

// multicast clusters of two pairs (see the kernel): PCG_GEMM_MC=0 / 1, pcg_gemm_set_variant(2) forces them on
bool g_mc_enabled = []() {
    const char* e = getenv("PCG_GEMM_MC");
    return e != nullptr && e[0] == '1';
}();

bool g_pair_enabled = []() {
    const char* e = getenv("PCG_GEMM_PAIR");
    return !(e != nullptr && e[0] == '0');
}();

}  // namespace

int gemm_bf16_impl(int mode, int act, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                   const float* bias, const void* aux, void* out, void* out2, int ldo, int force_bn,
                   cudaStream_t stream) {
    PCG_CHECK_ARG(M > 0 && N > 0 && K > 0, "pcg_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
    PCG_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "pcg_gemm_bf16: K, lda, ldb must be multiples of 8");
    PCG_CHECK_ARG(N % 32 == 0 && ldo % 8 == 0, "pcg_gemm_bf16: N %% 32 and ldo %% 8 must be 0 (N=%d ldo=%d)", N, ldo);
    PCG_CHECK_ARG(A && B && out, "pcg_gemm_bf16: null operand");
    PCG_CHECK_ARG((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
                   reinterpret_cast<uintptr_t>(out)) % 16 == 0, "pcg_gemm_bf16: operands must be 16-byte aligned");
    PCG_CHECK_ARG(mode != PCG_GEMM_BIAS_ACT || out2, "pcg_gemm_bf16: BIAS_ACT needs out2");
    PCG_CHECK_ARG((mode != PCG_GEMM_RESID_F32 && mode != PCG_GEMM_DACT) || aux, "pcg_gemm_bf16: mode needs aux");
    // CTA pairs (256 x 256 tiles) when the problem fills the 74 pairs with full-width tiles; force_bn 512 forces
    // them, any other force_bn the single-CTA kernel
    float single_cost = 0.f;
    const int single_bn = choose_bn(M, N, &single_cost);
    const bool pair = force_bn ? (force_bn == 512 || force_bn == 1024)
                               : (g_pair_enabled && N % 256 == 0 && ceil_div(M, 256) * (N / 256) >= sm_count() / 2 &&
                                  pair_cost(M, N) <= single_cost);
    const int bn = pair ? 256 : (force_bn ? force_bn : single_bn);
    // two pairs per cluster sharing A: needs an even number of column tiles and at least two row tiles of work per pair
    const bool mc = pair && (g_mc_enabled || force_bn == 1024) && N % 512 == 0;
    CUtensorMap ma, mb;
    int rc = get_tensor_map(&ma, A, M, K, lda, mc ? BM / 2 : BM);
    if (rc) return rc;
    rc = get_tensor_map(&mb, B, N, K, ldb, pair ? 128 : bn);
    if (rc) return rc;
    CUtensorMap mo = ma, mo2 = ma;  // placeholders unless the bulk-store epilogue is compiled in and applies
    if (PCG_GEMM_TS && (mode == PCG_GEMM_BF16 || mode == PCG_GEMM_BIAS_ACT)) {
        rc = get_store_map(&mo, out, M, N, ldo);
        if (rc) return rc;
        if (mode == PCG_GEMM_BIAS_ACT) {
            rc = get_store_map(&mo2, out2, M, N, ldo);
            if (rc) return rc;
        }
    }
    GemmParams p{M, N, K, bias, aux, out, out2, ldo, act};
    ProfileScope prof(PCG_PROF_GEMM, 2.0 * M * N * K, stream);
    if (mc) return dispatch_mode<256, 2, true>(mode, ma, mb, mo, mo2, p, stream);
    if (pair) return dispatch_mode<256, 2>(mode, ma, mb, mo, mo2, p, stream);
    switch (bn) {
        case 256: return dispatch_mode<256>(mode, ma, mb, mo, mo2, p, stream);
        case 192: return dispatch_mode<192>(mode, ma, mb, mo, mo2, p, stream);
        case 128: return dispatch_mode<128>(mode, ma, mb, mo, mo2, p, stream);
        case 64: return dispatch_mode<64>(mode, ma, mb, mo, mo2, p, stream);
        default: return set_error(-1, "pcg_gemm_bf16: unsupported BN %d", bn);
    }
}

}  // namespace pcg

extern "C" int pcg_gemm_bf16(int mode, int act, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                             const float* bias, const void* aux, void* out, void* out2, int ldo, void* stream) {
    return pcg::gemm_bf16_impl(mode, act, M, N, K, A, lda, B, ldb, bias, aux, out, out2, ldo, 0,
                               static_cast<cudaStream_t>(stream));
}
// tools / tests: 0 = single-CTA kernels only, 1 = CTA pairs allowed, 2 = CTA pairs in multicast clusters of four
extern "C" int pcg_gemm_set_variant(int v) {
    pcg::g_pair_enabled = v != 0;
    pcg::g_mc_enabled = v == 2;
    return 0;
}
// test hook: force the N tile width (64/128/192/256)
extern "C" int pcg_gemm_bf16_bn(int bn, int mode, int act, int M, int N, int K, const void* A, int lda, const void* B,
                                int ldb, const float* bias, const void* aux, void* out, void* out2, int ldo,
                                void* stream) {
    return pcg::gemm_bf16_impl(mode, act, M, N, K, A, lda, B, ldb, bias, aux, out, out2, ldo, bn,
                               static_cast<cudaStream_t>(stream));
}
