// C-ABI glue: error reporting, buffer sizing, and the host-side sequencer that runs the whole guidance step
// (sampler -> patch-embed GEMM -> L transformer blocks -> head/loss, and the dgrad-only backward to the image)
// as one native call per direction, so Python pays one FFI crossing per forward/backward.
//
// Replaces the eager op sequence of CLIP.forward (perceptor/losses/clip/clip.py:89-99) ->
// OpenCLIP.encode_images (perceptor/models/open_clip.py:109-123) -> VisionTransformer.forward
// (perceptor/models/ruclip/model.py:105-131) and the autograd tape behind it.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "pcg_common.cuh"

namespace pcg {

char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
static thread_local int g_launches = 0;
void count_launch() { ++g_launches; }
void reset_launch_count() { g_launches = 0; }
int launch_count() { return g_launches; }
int sm_count() {
    static int sms = []() {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        return n;
    }();
    return sms;
}

// ---- optional device-side timing of kernel families ----------------------------------------------------
struct ProfRecord {
    int kind;
    double work;
    cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;
void profile_begin(int kind, double work, cudaStream_t stream) {
    if (!g_prof_on) return;
    ProfRecord r{kind, work, nullptr, nullptr};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, stream);
    g_prof.push_back(r);
}
void profile_end(cudaStream_t stream) {
    if (!g_prof_on || g_prof.empty()) return;
    cudaEventRecord(g_prof.back().e1, stream);
}

namespace {

struct Bump {
    uint8_t* base;
    size_t off = 0;
    explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
    template <typename T>
    T* take(size_t count) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += align_up(count * sizeof(T), 256);
        return p;
    }
};

struct LayerStash {
    float* xb;     // [M,D]  residual after attention
    uint16_t* qkv; // [M,3Da]  (Da = heads * head stride: D for head dim 64)
    uint16_t* o;   // [M,Da]
    float* lse;    // [n*heads*T]
    uint16_t* h;   // [M,4D] act'(pre-activation) of the MLP (the derivative is stored, not h)
};

struct Stash {
    float* v;                 // [M,D] pre-ln_pre
    float* x0;                // [M,D] first residual; x(l+1) lives at x0 + (l+1)*M*D stride-aligned
    size_t x_stride;          // elements between consecutive x(l)
    LayerStash layer0;        // first layer's buffers
    size_t layer_stride;      // bytes between consecutive layers' stash blocks
    size_t total;
};

// when `layers_kept` == 1 every layer aliases the same block (forward-only mode)
Stash carve_stash(void* base, const pcg_vit_config& c, int n, int layers_kept) {
    const size_t M = static_cast<size_t>(n) * c.tokens, D = c.width;
    const size_t Da = static_cast<size_t>(c.heads) * pcg_head_stride(c.head_dim);  // attention-side width (padded heads)
    Bump b(base);
    Stash s;
    s.v = b.take<float>(M * D);
    s.x_stride = align_up(M * D * sizeof(float), 256) / sizeof(float);
    const int nx = (layers_kept == 1) ? 2 : c.layers + 1;
    s.x0 = b.take<float>(s.x_stride * nx);
    const size_t before = b.off;
    s.layer0.xb = b.take<float>(M * D);
    s.layer0.qkv = b.take<uint16_t>(M * 3 * Da);
    s.layer0.o = b.take<uint16_t>(M * Da);
    s.layer0.lse = b.take<float>(static_cast<size_t>(n) * c.heads * c.tokens);
    s.layer0.h = b.take<uint16_t>(M * c.mlp);
    s.layer_stride = b.off - before;
    b.off = before + s.layer_stride * layers_kept;
    s.total = b.off;
    return s;
}

LayerStash layer_stash(const Stash& s, int l, bool kept) {
    if (!kept) return s.layer0;
    const size_t d = s.layer_stride * l;
    LayerStash r;
    r.xb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s.layer0.xb) + d);
    r.qkv = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(s.layer0.qkv) + d);
    r.o = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(s.layer0.o) + d);
    r.lse = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s.layer0.lse) + d);
    r.h = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(s.layer0.h) + d);
    return r;
}
float* stash_x(const Stash& s, int l, bool kept) { return s.x0 + s.x_stride * (kept ? l : (l & 1)); }

struct Work {
    uint16_t* patches;  // [P,kpad]   fwd: sampler output; bwd: d_patches
    float* patch_out;   // [P,D] f32  fwd: conv1 output;   bwd: d_patch bf16 [P,D] (aliases)
    uint16_t* y;        // [M,D]      LN output / dy
    uint16_t* a;        // [M,4D]     activation output / dH
    uint16_t* dxb;      // [M,D]      the residual-stream gradient (bf16, accumulated in place by the LN backward)
    uint16_t* dqkv;     // [M,3D]
    uint16_t* d_o;      // [M,D]
    float* delta;       // pcg_attn_bwd_workspace_bytes: [n*heads*T] (+ f32 dQ sums [M,D] for long sequences)
    float* head_ws;     // pcg_head_workspace_bytes
    void* nostash;      // forward-only stash region
    size_t total;
};

Work carve_work(void* base, const pcg_vit_config& c, int n) {
    const size_t M = static_cast<size_t>(n) * c.tokens, D = c.width, P = static_cast<size_t>(n) * c.grid * c.grid;
    const size_t Da = static_cast<size_t>(c.heads) * pcg_head_stride(c.head_dim);
    Bump b(base);
    Work w;
    w.patches = b.take<uint16_t>(P * c.kpad);
    w.patch_out = b.take<float>(P * D);
    w.y = b.take<uint16_t>(M * D);
    w.a = b.take<uint16_t>(M * c.mlp);
    w.dxb = b.take<uint16_t>(M * D);
    w.dqkv = b.take<uint16_t>(M * 3 * Da);
    w.d_o = b.take<uint16_t>(M * Da);
    w.delta = b.take<float>(pcg_attn_bwd_workspace_bytes(n, c.tokens, c.heads) / sizeof(float));
    w.head_ws = b.take<float>(pcg_head_workspace_bytes(n, c.width, c.embed) / sizeof(float));
    w.nostash = base ? static_cast<uint8_t*>(base) + b.off : nullptr;
    b.off += carve_stash(nullptr, c, n, 1).total;
    w.total = b.off;
    return w;
}

int check_cfg(const pcg_vit_config* c) {
    PCG_CHECK_ARG(c != nullptr, "null config");
    PCG_CHECK_ARG(c->width > 0 && c->width % 128 == 0 && c->heads > 0 && c->heads * c->head_dim == c->width &&
                      (c->head_dim == 64 || (c->head_dim > 64 && c->head_dim <= 128 && c->head_dim % 8 == 0)),
                  "width %d / heads %d / head dim %d: need heads * head_dim == width, width a multiple of 128 and a head "
                  "dim of 64 or 72..128 (padded to 128)", c->width, c->heads, c->head_dim);
    PCG_CHECK_ARG(c->grid * c->patch == c->image_size && c->tokens == c->grid * c->grid + 1, "inconsistent grid/tokens");
    PCG_CHECK_ARG(c->kpatch == 3 * c->patch * c->patch && c->kpad >= c->kpatch && c->kpad % 64 == 0, "bad kpatch/kpad");
    PCG_CHECK_ARG(c->mlp % 128 == 0 && c->layers > 0 && c->embed > 0, "bad mlp/layers/embed");
    return 0;
}

int check_args(const pcg_guidance_args* a, bool bwd) {
    PCG_CHECK_ARG(a && a->cfg && a->w && a->w->layers_host, "pcg_guidance: null args");
    if (int rc = check_cfg(a->cfg)) return rc;
    PCG_CHECK_ARG(a->images || bwd, "pcg_guidance: images is null");
    PCG_CHECK_ARG(a->cuts && a->tabs && a->n_cut > 0 && a->max_in_w > 0, "pcg_guidance: cutout table missing");
    PCG_CHECK_ARG(a->workspace && a->workspace_bytes >= pcg_workspace_bytes(a->cfg, a->n_cut),
                  "pcg_guidance: workspace too small (%zu < %zu)", a->workspace_bytes,
                  pcg_workspace_bytes(a->cfg, a->n_cut));
    if (a->want_grad || bwd)
        PCG_CHECK_ARG(a->stash && a->stash_bytes >= pcg_stash_bytes(a->cfg, a->n_cut),
                      "pcg_guidance: stash too small (%zu < %zu)", a->stash_bytes, pcg_stash_bytes(a->cfg, a->n_cut));
    if (bwd)
        PCG_CHECK_ARG(a->d_images && (a->d_enc || (a->targets && a->tweights && a->n_targets > 0)),
                      "pcg_guidance_bwd: missing d_images, or neither targets nor d_enc given");
    return 0;
}

// Last block on the class-token rows only (see pcg_set_pooled_last_block in pcg.h)
bool g_pooled_last = []() {
    const char* e = getenv("PCG_FULL_LAST_BLOCK");
    return !(e != nullptr && e[0] == '1');
}();

#define PCG_TRY(expr)          \
    do {                       \
        int _rc = (expr);      \
        if (_rc) return _rc;   \
    } while (0)

}  // namespace
}  // namespace pcg

using namespace pcg;

extern "C" const char* pcg_last_error(void) { return last_error_buf(); }
extern "C" int pcg_abi_version(void) { return PCG_ABI_VERSION; }
extern "C" int pcg_device_sm_count(void) { return sm_count(); }
extern "C" int pcg_head_stride(int head_dim) { return head_dim == 64 ? 64 : 128; }
extern "C" int pcg_last_launch_count(void) { return launch_count(); }
extern "C" int pcg_set_pooled_last_block(int on) {
    const int prev = g_pooled_last ? 1 : 0;
    g_pooled_last = on != 0;
    return prev;
}

extern "C" int pcg_profile_enable(int on) {
    g_prof_on = on != 0;
    return 0;
}
// Sums the records since the last collect into ms[kind], work[kind], count[kind] (arrays of PCG_PROF_KINDS).
extern "C" int pcg_profile_collect(double* ms, double* work, int* count) {
    PCG_CHECK_ARG(ms && work && count, "pcg_profile_collect: null output");
    for (int i = 0; i < PCG_PROF_KINDS; ++i) ms[i] = work[i] = 0.0, count[i] = 0;
    for (ProfRecord& r : g_prof) {
        float t = 0.f;
        if (r.e0 && r.e1 && cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess &&
            r.kind >= 0 && r.kind < PCG_PROF_KINDS) {
            ms[r.kind] += t;
            work[r.kind] += r.work;
            count[r.kind] += 1;
        }
        if (r.e0) cudaEventDestroy(r.e0);
        if (r.e1) cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    return 0;
}

extern "C" size_t pcg_workspace_bytes(const pcg_vit_config* cfg, int n_cut) {
    if (cfg == nullptr || n_cut <= 0) return 0;
    return carve_work(nullptr, *cfg, n_cut).total;
}
extern "C" size_t pcg_stash_bytes(const pcg_vit_config* cfg, int n_cut) {
    if (cfg == nullptr || n_cut <= 0) return 0;
    return carve_stash(nullptr, *cfg, n_cut, cfg->layers).total;
}

extern "C" int pcg_guidance_fwd(const pcg_guidance_args* a, void* stream) {
    PCG_TRY(check_args(a, false));
    reset_launch_count();
    const pcg_vit_config& c = *a->cfg;
    const pcg_vit_weights& w = *a->w;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = a->n_cut, T = c.tokens, D = c.width;
    const int M = n * T, P = n * c.grid * c.grid;
    const bool wide = c.head_dim != 64;                      // ViT-H/14, ViT-g/14: heads padded to 128 columns
    const int Da = c.heads * pcg_head_stride(c.head_dim);    // attention-side width
    const bool keep = a->want_grad != 0;
    Work wk = carve_work(a->workspace, c, n);
    Stash st = keep ? carve_stash(a->stash, c, n, c.layers) : carve_stash(wk.nostash, c, n, 1);

    if (c.kpad != c.kpatch) PCG_CUDA(cudaMemsetAsync(wk.patches, 0, static_cast<size_t>(P) * c.kpad * 2, s));
    PCG_TRY(pcg_sampler_fwd(a->images, a->B, a->H, a->W, a->cuts, n, a->tabs, c.image_size, c.patch, c.kpad,
                            a->mean_host, a->std_host, wk.patches, nullptr, a->max_in_w, stream));
    PCG_TRY(pcg_gemm_bf16(PCG_GEMM_F32, c.act, P, D, c.kpad, wk.patches, c.kpad, w.conv1, c.kpad, nullptr, nullptr,
                          wk.patch_out, nullptr, D, stream));
    PCG_TRY(pcg_embed_fwd(wk.patch_out, w.cls, w.pos, w.ln_pre_g, w.ln_pre_b, st.v, stash_x(st, 0, keep), n, T, D,
                          stream));
    const int hs = pcg_head_stride(c.head_dim);
    for (int l = 0; l < c.layers; ++l) {
        const pcg_layer_weights& lw = w.layers_host[l];
        const LayerStash ls = layer_stash(st, l, keep);
        float* x_in = stash_x(st, l, keep);
        float* x_out = stash_x(st, l + 1, keep);
        // The head reads only the class-token row of the last block's output: there, K and V are projected for every
        // row but q, the attention output, the out-projection, ln_2 and the MLP run on the n class-token rows
        // (row stride T in the same buffers, so the backward finds everything where the full block would put it).
        const bool pooled = g_pooled_last && l == c.layers - 1 && T > 1;
        const int Mr = pooled ? n : M;          // rows of the row-wise ops after the K / V projection
        const int rs = pooled ? T : 1;          // their row step
        PCG_TRY(pcg_layernorm_fwd(x_in, lw.ln1_g, lw.ln1_b, wk.y, M, D, stream));
        if (!pooled) {
            PCG_TRY(pcg_gemm_bf16(PCG_GEMM_BF16, c.act, M, 3 * Da, D, wk.y, D, lw.w_qkv, D, lw.b_qkv, nullptr, ls.qkv,
                                  nullptr, 3 * Da, stream));
            PCG_TRY(wide ? pcg_attn_fwd_wide(ls.qkv, ls.o, ls.lse, n, T, c.heads, stream)
                         : pcg_attn_fwd(ls.qkv, ls.o, ls.lse, n, T, c.heads, stream));
        } else {
            const uint16_t* w_kv = static_cast<const uint16_t*>(lw.w_qkv) + static_cast<size_t>(Da) * D;
            PCG_TRY(pcg_gemm_bf16(PCG_GEMM_BF16, c.act, M, 2 * Da, D, wk.y, D, w_kv, D, lw.b_qkv + Da, nullptr, ls.qkv + Da,
                                  nullptr, 3 * Da, stream));
            PCG_TRY(pcg_gemm_bf16(PCG_GEMM_BF16, c.act, n, Da, D, wk.y, T * D, lw.w_qkv, D, lw.b_qkv, nullptr, ls.qkv,
                                  nullptr, T * 3 * Da, stream));
            PCG_TRY(pcg_attn_cls_fwd(ls.qkv, ls.o, ls.lse, n, T, c.heads, hs, stream));
        }
        PCG_TRY(pcg_gemm_bf16(PCG_GEMM_RESID_F32, c.act, Mr, D, Da, ls.o, rs * Da, lw.w_out, Da, lw.b_out, x_in, ls.xb,
                              nullptr, rs * D, stream));
        PCG_TRY(pcg_layernorm_fwd_rows(ls.xb, lw.ln2_g, lw.ln2_b, wk.y, Mr, D, rs, stream));
        PCG_TRY(pcg_gemm_bf16(PCG_GEMM_BIAS_ACT, c.act, Mr, c.mlp, D, wk.y, rs * D, lw.w_fc, D, lw.b_fc, nullptr, ls.h, wk.a,
                              rs * c.mlp, stream));
        PCG_TRY(pcg_gemm_bf16(PCG_GEMM_RESID_F32, c.act, Mr, D, c.mlp, wk.a, rs * c.mlp, lw.w_proj, c.mlp, lw.b_proj, ls.xb,
                              x_out, nullptr, rs * D, stream));
    }
    PCG_TRY(pcg_head_loss(stash_x(st, c.layers, keep), w.ln_post_g, w.ln_post_b, w.proj, a->targets, a->tweights, n, T, D,
                          c.embed, a->targets ? a->n_targets : 0, a->loss_scale, a->normalize, a->loss_sum, a->enc_out,
                          nullptr, nullptr, nullptr, wk.head_ws, stream));
    return 0;
}

extern "C" int pcg_guidance_bwd(const pcg_guidance_args* a, void* stream) {
    PCG_TRY(check_args(a, true));
    reset_launch_count();
    const pcg_vit_config& c = *a->cfg;
    const pcg_vit_weights& w = *a->w;
    const int n = a->n_cut, T = c.tokens, D = c.width;
    const int M = n * T, P = n * c.grid * c.grid;
    const bool wide = c.head_dim != 64;
    const int Da = c.heads * pcg_head_stride(c.head_dim);
    Work wk = carve_work(a->workspace, c, n);
    Stash st = carve_stash(a->stash, c, n, c.layers);

    // head: recompute the (tiny) forward of the head and emit d(loss)/dx at the class-token rows
    PCG_TRY(pcg_head_loss(stash_x(st, c.layers, true), w.ln_post_g, w.ln_post_b, w.proj, a->targets, a->tweights, n, T, D,
                          c.embed, a->targets ? a->n_targets : 0, a->d_enc ? 1.0f : a->loss_scale, a->normalize, nullptr,
                          nullptr, a->d_enc, nullptr, wk.dxb, wk.head_ws, stream));
    const int hs = pcg_head_stride(c.head_dim);
    for (int l = c.layers - 1; l >= 0; --l) {
        const pcg_layer_weights& lw = w.layers_host[l];
        const LayerStash ls = layer_stash(st, l, true);
        // last block: the gradient arrives on the class-token rows only (the head zeroes the others), and the forward
        // kept only those rows of x_mid / act' / o -- the same row-stepped calls in reverse
        const bool pooled = g_pooled_last && l == c.layers - 1 && T > 1;
        const int Mr = pooled ? n : M;
        const int rs = pooled ? T : 1;
        // MLP: dH = (dx W_proj) * act'(h) ; dy = dH W_fc ; dx += ln_2'(dy)
        PCG_TRY(pcg_gemm_bf16(PCG_GEMM_DACT, c.act, Mr, c.mlp, D, wk.dxb, rs * D, lw.w_proj_t, D, nullptr, ls.h, wk.a,
                              nullptr, rs * c.mlp, stream));
        PCG_TRY(pcg_gemm_bf16(PCG_GEMM_BF16, c.act, Mr, D, c.mlp, wk.a, rs * c.mlp, lw.w_fc_t, c.mlp, nullptr, nullptr, wk.y,
                              nullptr, rs * D, stream));
        PCG_TRY(pcg_layernorm_bwd_rows(wk.y, ls.xb, lw.ln2_g, nullptr, wk.dxb, Mr, D, rs, stream));
        // attention: dO = dx W_out ; dqkv = attn'(dO) ; dy = dqkv W_qkv ; dx += ln_1'(dy)
        PCG_TRY(pcg_gemm_bf16(PCG_GEMM_BF16, c.act, Mr, Da, D, wk.dxb, rs * D, lw.w_out_t, D, nullptr, nullptr, wk.d_o,
                              nullptr, rs * Da, stream));
        if (pooled)
            PCG_TRY(pcg_attn_cls_bwd(ls.qkv, ls.o, wk.d_o, ls.lse, wk.dqkv, n, T, c.heads, hs, stream));
        else
            PCG_TRY(wide ? pcg_attn_bwd_wide(ls.qkv, ls.o, wk.d_o, ls.lse, wk.delta, wk.dqkv, n, T, c.heads, stream)
                         : pcg_attn_bwd(ls.qkv, ls.o, wk.d_o, ls.lse, wk.delta, wk.dqkv, n, T, c.heads, stream));
        PCG_TRY(pcg_gemm_bf16(PCG_GEMM_BF16, c.act, M, D, 3 * Da, wk.dqkv, 3 * Da, lw.w_qkv_t, 3 * Da, nullptr, nullptr,
                              wk.y, nullptr, D, stream));
        PCG_TRY(pcg_layernorm_bwd(wk.y, stash_x(st, l, true), lw.ln1_g, nullptr, wk.dxb, M, D, stream));
    }
    uint16_t* d_patch = reinterpret_cast<uint16_t*>(wk.patch_out);  // bf16 [P,D] fits in the f32 [P,D] buffer
    PCG_TRY(pcg_embed_bwd(nullptr, wk.dxb, st.v, w.ln_pre_g, d_patch, n, T, D, stream));
    PCG_TRY(pcg_gemm_bf16(PCG_GEMM_BF16, c.act, P, c.kpad, D, d_patch, D, w.conv1_t, D, nullptr, nullptr, wk.patches,
                          nullptr, c.kpad, stream));
    PCG_TRY(pcg_sampler_bwd(wk.patches, nullptr, a->B, a->H, a->W, a->cuts, n, a->tabs, c.image_size, c.patch, c.kpad,
                            a->std_host, a->d_images, a->max_in_w, stream));
    return 0;
}
