// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  out[M,N] = epilogue( A[M,K] * B[N,K]^T ).
//
//   * A and B are bf16, K-major (row-major [rows, K]) — exactly nn.Linear's activation / weight layout, so the
//     forward GEMMs consume weights as stored and the dgrad GEMMs consume pre-transposed copies.
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) stages 128 x 64 (A) and BN x 64 (B) tiles through an
//     mbarrier ring; one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered
//     fp32 accumulator in tensor memory; eight epilogue warps drain it with tcgen05.ld and apply the fused
//     epilogue (bias / QuickGELU / residual add / activation derivative) straight from registers.
//   * grid = min(#tiles, #SMs); each CTA walks tiles t = blockIdx.x + i*gridDim.x (n fastest so that the CTAs
//     running concurrently share A row-panels in L2; B (weights) is L2 resident).
//
// Replaces: nn.Linear forward/dgrad and conv1-as-GEMM in the reference ViT (perceptor/models/ruclip/model.py:
// 31-39, 43-49, 85-91), which today are eager cuBLAS calls plus separate bias/activation/residual kernels.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>
#include <unordered_map>

#include "pcg_common.cuh"
#include "pcg_ptx.cuh"

namespace pcg {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kCtrlWarps = 4;  // warp 0: TMA producer, warp 1: MMA issuer, warp 2: TMEM allocator, warp 3: spare
constexpr int kEpiWarps = 8;   // two warpgroups; warp%4 selects the TMEM lane quarter, warpgroup the column half
constexpr int kThreads = 32 * (kCtrlWarps + kEpiWarps);
constexpr int kABytes = BM * BK * 2;

template <int BN>
struct TileCfg {
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStages = (BN == 256) ? 4 : (BN == 192) ? 5 : (BN == 128) ? 6 : 8;
    static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int kBarBytes = 256;
    static constexpr int kSmem = kStages * (kABytes + kBBytes) + kBarBytes + 1024;  // +1024: manual alignment
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

__device__ __forceinline__ float act_fwd(float h, int act) {
    if (act == PCG_ACT_QUICKGELU) return __fdividef(h, 1.0f + __expf(-1.702f * h));
    return 0.5f * h * (1.0f + erff(h * 0.70710678118654752f));
}
__device__ __forceinline__ float act_bwd(float h, int act) {
    if (act == PCG_ACT_QUICKGELU) {
        const float s = __fdividef(1.0f, 1.0f + __expf(-1.702f * h));
        return s * (1.0f + 1.702f * h * (1.0f - s));
    }
    return 0.5f * (1.0f + erff(h * 0.70710678118654752f)) + h * 0.3989422804014327f * __expf(-0.5f * h * h);
}

template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p) {
    using Cfg = TileCfg<BN>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kBBytes = Cfg::kBBytes;

    extern __shared__ uint8_t smem_raw[];
    // 128-byte swizzle atoms need 1024-byte aligned stage buffers.
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBBytes));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tfull_bar = bars + 2 * kStages;
    uint64_t* tempty_bar = bars + 2 * kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (p.M + BM - 1) / BM;
    const int num_n = (p.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / num_n) * BM;
                const int n0 = (tile % num_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], kABytes + kBBytes);
                    tma_load_2d(&map_a, &full_bar[stage], smem_a + stage * kABytes, kb * BK, m0, kEvictNormal);
                    tma_load_2d(&map_b, &full_bar[stage], smem_b + stage * kBBytes, kb * BK, n0, kEvictLast);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
                        umma_f16(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= kCtrlWarps) {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;             // TMEM lanes [32*quarter, +32) are the ones this warp may read
        const int half = (warp - kCtrlWarps) >> 2;  // column half of the tile
        constexpr int kHalfN = BN / 2;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / num_n) * BM;
            const int n0 = (tile % num_n) * BN;
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const int grow = m0 + quarter * 32 + lane;
            const bool row_ok = grow < p.M;
            const size_t row_off = static_cast<size_t>(grow) * p.ldo;
#pragma unroll 1
            for (int c = 0; c < kHalfN; c += 32) {
                const int col0 = n0 + half * kHalfN + c;
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * kHalfN + c,
                              r);
                tmem_wait_ld();
                if (row_ok && col0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                            v[j] += b.x;
                            v[j + 1] += b.y;
                            v[j + 2] += b.z;
                            v[j + 3] += b.w;
                        }
                    }
                    if constexpr (MODE == PCG_GEMM_BF16) {
                        uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + row_off + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                                                pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                    } else if constexpr (MODE == PCG_GEMM_BIAS_ACT) {
                        uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + row_off + col0);
                        uint4* dst2 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out2) + row_off + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                                                pack_bf16(v[8 * j + 6], v[8 * j + 7]));
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], p.act);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst2[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                 pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                                                 pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                    } else if constexpr (MODE == PCG_GEMM_RESID_F32) {
                        const float4* res = reinterpret_cast<const float4*>(static_cast<const float*>(p.aux) + row_off + col0);
                        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + row_off + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 x = res[j];
                            dst[j] = make_float4(x.x + v[4 * j], x.y + v[4 * j + 1], x.z + v[4 * j + 2],
                                                 x.w + v[4 * j + 3]);
                        }
                    } else if constexpr (MODE == PCG_GEMM_DACT) {
                        const uint4* hp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.aux) + row_off + col0);
                        uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + row_off + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 hv = hp[j];
                            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&hw[q]);
                                v[8 * j + 2 * q] *= act_bwd(__low2float(h2), p.act);
                                v[8 * j + 2 * q + 1] *= act_bwd(__high2float(h2), p.act);
                            }
                            dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                                                pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                        }
                    } else {  // PCG_GEMM_F32
                        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + row_off + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
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

template <int BN, int MODE>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t stream) {
    using Cfg = TileCfg<BN>;
    static bool configured = false;
    if (!configured) {
        PCG_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmem));
        configured = true;
    }
    const int tiles = ceil_div(p.M, BM) * ceil_div(p.N, BN);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    gemm_tcgen05_kernel<BN, MODE><<<grid, kThreads, Cfg::kSmem, stream>>>(ma, mb, p);
    PCG_LAUNCH_CHECK("gemm_tcgen05_kernel");
    return 0;
}

template <int BN>
int dispatch_mode(int mode, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t s) {
    switch (mode) {
        case PCG_GEMM_BF16: return launch_gemm<BN, PCG_GEMM_BF16>(ma, mb, p, s);
        case PCG_GEMM_BIAS_ACT: return launch_gemm<BN, PCG_GEMM_BIAS_ACT>(ma, mb, p, s);
        case PCG_GEMM_RESID_F32: return launch_gemm<BN, PCG_GEMM_RESID_F32>(ma, mb, p, s);
        case PCG_GEMM_DACT: return launch_gemm<BN, PCG_GEMM_DACT>(ma, mb, p, s);
        case PCG_GEMM_F32: return launch_gemm<BN, PCG_GEMM_F32>(ma, mb, p, s);
        default: return set_error(-1, "pcg_gemm_bf16: unknown mode %d", mode);
    }
}

// Pick the N tile that minimises (waves x per-tile cost).  Per-tile MMA time is proportional to BN; narrow
// tiles pay more shared-memory traffic per flop, hence the mild penalty.
int choose_bn(int M, int N) {
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
    return best;
}

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
    const int bn = force_bn ? force_bn : choose_bn(M, N);
    CUtensorMap ma, mb;
    int rc = get_tensor_map(&ma, A, M, K, lda, BM);
    if (rc) return rc;
    rc = get_tensor_map(&mb, B, N, K, ldb, bn);
    if (rc) return rc;
    GemmParams p{M, N, K, bias, aux, out, out2, ldo, act};
    ProfileScope prof(PCG_PROF_GEMM, 2.0 * M * N * K, stream);
    switch (bn) {
        case 256: return dispatch_mode<256>(mode, ma, mb, p, stream);
        case 192: return dispatch_mode<192>(mode, ma, mb, p, stream);
        case 128: return dispatch_mode<128>(mode, ma, mb, p, stream);
        case 64: return dispatch_mode<64>(mode, ma, mb, p, stream);
        default: return set_error(-1, "pcg_gemm_bf16: unsupported BN %d", bn);
    }
}

}  // namespace pcg

extern "C" int pcg_gemm_bf16(int mode, int act, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                             const float* bias, const void* aux, void* out, void* out2, int ldo, void* stream) {
    return pcg::gemm_bf16_impl(mode, act, M, N, K, A, lda, B, ldb, bias, aux, out, out2, ldo, 0,
                               static_cast<cudaStream_t>(stream));
}
// test hook: force the N tile width (64/128/192/256)
extern "C" int pcg_gemm_bf16_bn(int bn, int mode, int act, int M, int N, int K, const void* A, int lda, const void* B,
                                int ldb, const float* bias, const void* aux, void* out, void* out2, int ldo,
                                void* stream) {
    return pcg::gemm_bf16_impl(mode, act, M, N, K, A, lda, B, ldb, bias, aux, out, out2, ldo, bn,
                               static_cast<cudaStream_t>(stream));
}
