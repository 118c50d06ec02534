/* pcg.h — C ABI of libpcg.so: the B200-native CLIP-guidance hot path (loss forward + backward to the image).
 *
 * This is the drop-in boundary for perceptor's one data-parallel hot path.  The reference has no FFI on this
 * path (it is eager PyTorch); each entry point below names the reference code it replaces (paths relative to
 * the reference repo root, perceptor v0.6.7).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`.
 *  - the library never allocates or frees device memory; the caller sizes buffers with pcg_workspace_bytes()
 *    and pcg_stash_bytes() and owns them.
 *  - `stream` is a cudaStream_t passed as void*; every launch goes to it; no host synchronisation inside.
 *  - return value: 0 = ok, negative = argument error, positive = cudaError_t.  pcg_last_error() returns a
 *    thread-local, human-readable message for the last non-zero return.
 *  - bf16 buffers are passed as void* (raw __nv_bfloat16 bits).
 */
#ifndef PCG_H_
#define PCG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCG_ABI_VERSION 4

enum { PCG_ACT_QUICKGELU = 0, PCG_ACT_GELU = 1 };

/* GEMM epilogues (C = A[M,K] * B[N,K]^T, bf16 inputs, fp32 accumulation in tensor memory). */
enum {
    PCG_GEMM_BF16 = 0,     /* out(bf16) = acc (+ bias)                                     */
    PCG_GEMM_BIAS_ACT = 1, /* h = acc + bias ; out(bf16) = act'(h) ; out2(bf16) = act(h)   */
    PCG_GEMM_RESID_F32 = 2,/* out(f32)  = aux(f32 residual) + acc + bias                   */
    PCG_GEMM_DACT = 3,     /* out(bf16) = acc * aux(bf16: the act'(h) stored by BIAS_ACT)  */
    PCG_GEMM_F32 = 4       /* out(f32)  = acc (+ bias)                                     */
};

/* Vision-transformer shape.  Mirrors the constructor arguments of the in-tree OpenAI-CLIP VisionTransformer
 * (perceptor/models/ruclip/model.py:72-103) that open_clip builds for perceptor/models/open_clip.py:65-72. */
typedef struct pcg_vit_config {
    int32_t image_size; /* R: 224 / 336                                   */
    int32_t patch;      /* p: 32 / 16 / 14                                */
    int32_t grid;       /* g = R / p                                      */
    int32_t tokens;     /* T = g*g + 1                                    */
    int32_t width;      /* D                                              */
    int32_t layers;     /* L                                              */
    int32_t heads;      /* D / head_dim                                   */
    int32_t mlp;        /* 4*D                                            */
    int32_t embed;      /* E: output dim of `proj`                        */
    int32_t kpatch;     /* 3*p*p                                          */
    int32_t kpad;       /* kpatch rounded up to a multiple of 64          */
    int32_t act;        /* PCG_ACT_*                                      */
    int32_t head_dim;   /* 64 (tcgen05 attention), or 80 / 88: ViT-H/14 and  */
                        /* ViT-g/14, stored padded to 128 columns per head    */
} pcg_vit_config;

/* Columns one head occupies in qkv / attention-output buffers: head_dim 64 -> 64, wider heads -> 128 (zero padded by
 * the weight packing).  The attention-side width of a tower is heads * pcg_head_stride(head_dim). */
int pcg_head_stride(int head_dim);

/* One ResidualAttentionBlock (ruclip/model.py:25-54).  w_* are bf16 [out,in] row-major (nn.Linear layout);
 * w_*_t are their transposes [in,out] used by the dgrad GEMMs.  The 1/sqrt(64) attention scale is folded
 * into the q rows of w_qkv / b_qkv (exact: a power of two).  With head_dim != 64 every head of q, k, v occupies 128
 * rows of w_qkv (80 / 88 real + zero rows) and 128 columns of w_out: 3D and D below read 3*Da and Da = heads*128
 * on the attention side. */
typedef struct pcg_layer_weights {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    const void *w_qkv, *w_qkv_t; /* [3D,D], [D,3D] */
    const void *w_out, *w_out_t; /* [D,D]          */
    const void *w_fc, *w_fc_t;   /* [4D,D], [D,4D] */
    const void *w_proj, *w_proj_t; /* [D,4D], [4D,D] */
    const float *b_qkv, *b_out, *b_fc, *b_proj;
} pcg_layer_weights;

typedef struct pcg_vit_weights {
    const void *conv1, *conv1_t;      /* bf16 [D,kpad], [kpad,D]; columns >= kpatch are zero   */
    const float *cls, *pos;           /* [D], [T,D]                                            */
    const float *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;
    const float *proj;                /* fp32 [D,E]                                            */
    const pcg_layer_weights *layers_host; /* HOST array of `layers` entries                    */
} pcg_vit_weights;

/* Resize tables (S1).  One entry per distinct (input size, method); built on the host with the reference's own
 * fp32 arithmetic (perceptor/transforms/resize/resize_right.py:192-285) and uploaded once.
 *   desc[id] = {taps, left_off, weight_off, inv_off, in_size, 0,0,0}   (int32 x 8)
 *   left[left_off + o]            = first input index of output o (may be <0 or reach >= in_size: zero pad)
 *   weight[weight_off + o*taps+t] = weight of tap t of output o
 *   inv[inv_off + 2*i + {0,1}]    = [lo,hi) range of outputs that read input i                */
typedef struct pcg_resize_tables {
    const int32_t *desc;
    const int32_t *left;
    const float *weight;
    const int32_t *inv;
    int32_t n_desc;
} pcg_resize_tables;

/* Cutout rows (S0), int32 x 8 per cutout: {b, y0, x0, size_h, size_w, table_id_h, table_id_w, 0}. */
#define PCG_CUT_STRIDE 8

const char *pcg_last_error(void);
int pcg_abi_version(void);
int pcg_device_sm_count(void);

/* ---- S0+S1+S2: cutout + antialiased resize + normalize --------------------------------------------------
 * replaces resize(images, out_shape=image_size) + torchvision Normalize in
 * perceptor/models/open_clip.py:109-118 (resize = perceptor/transforms/resize/resize_right.py:34-189).
 * images f32 [B,3,H,W].  patches_bf16 (nullable) [n_cut*g*g, kpad]: row = (cutout, gy, gx),
 * col = c*p*p + py*p + px (the flattening of conv1.weight, ruclip/model.py:85-91), cols >= 3*p*p untouched
 * (caller zero-fills once).  out_f32 (nullable) [n_cut,3,R,R].  mean/std_inv: 3 host floats each. */
int pcg_sampler_fwd(const float *images, int B, int H, int W, const int32_t *cuts, int n_cut,
                    const pcg_resize_tables *tabs, int R, int patch, int kpad, const float *mean_host,
                    const float *std_host, void *patches_bf16, float *out_f32, int max_in_w, void *stream);
/* backward of the above: d_patches bf16 [n_cut*g*g, kpad] -> atomically accumulated into d_images f32
 * [B,3,H,W] (caller zeroes it).  replaces autograd through resize_right.apply_weights (:288-318). */
int pcg_sampler_bwd(const void *d_patches_bf16, const float *d_out_f32, int B, int H, int W, const int32_t *cuts,
                    int n_cut, const pcg_resize_tables *tabs, int R, int patch, int kpad, const float *std_host,
                    float *d_images, int max_in_w, void *stream);

/* ---- tcgen05 GEMM: out[M,N] = epilogue(A[M,K] * B[N,K]^T) -------------------------------------------------
 * replaces nn.Linear / conv1-as-GEMM forward and dgrad (ruclip/model.py:31-39,43-49,85-91).
 * A bf16 [M,lda], B bf16 [N,ldb], K % 8 == 0, lda/ldb % 8 == 0, N % 32 == 0; out/out2/aux have row stride ldo. */
int pcg_gemm_bf16(int mode, int act, int M, int N, int K, const void *A, int lda, const void *B, int ldb,
                  const float *bias, const void *aux, void *out, void *out2, int ldo, void *stream);

/* ---- LayerNorm (fp32 residual stream in, bf16 GEMM operand out) ------------------------------------------
 * replaces LayerNorm.forward (ruclip/model.py:11-17), eps = 1e-5. */
int pcg_layernorm_fwd(const float *x, const float *gamma, const float *beta, void *y_bf16, int rows, int D,
                      void *stream);
/* dx += LN'(x)^T dy.  dx_io != NULL: the residual gradient is fp32 (dx_io, read and written) and dx_bf16 receives
 * its bf16 copy (next dgrad GEMM operand).  dx_io == NULL: dx_bf16 IS the residual gradient, accumulated in place
 * (10 instead of 16 bytes per element; what pcg_guidance_bwd uses). */
int pcg_layernorm_bwd(const void *dy_bf16, const float *x, const float *gamma, float *dx_io, void *dx_bf16,
                      int rows, int D, void *stream);
/* The same two on every row_step-th row of their [*, D] buffers (row i of the call is row i*row_step of x, y, dy and
 * dx): the class-token rows (row_step = T) of the last block, see pcg_set_pooled_last_block. */
int pcg_layernorm_fwd_rows(const float *x, const float *gamma, const float *beta, void *y_bf16, int rows, int D,
                           int row_step, void *stream);
int pcg_layernorm_bwd_rows(const void *dy_bf16, const float *x, const float *gamma, float *dx_io, void *dx_bf16,
                           int rows, int D, int row_step, void *stream);

/* ---- class token + positional embedding + ln_pre (ruclip/model.py:109-120) ------------------------------
 * patch_out f32 [n*g*g, D] -> v f32 [n*T, D] (pre-LN, stashed) and x0 f32 [n*T, D]. */
int pcg_embed_fwd(const float *patch_out, const float *cls, const float *pos, const float *gamma,
                  const float *beta, float *v, float *x0, int n, int T, int D, void *stream);
/* dx0 [n*T,D], given as f32 (dx0) or bf16 (dx0_bf16), exactly one non-NULL -> d_patch bf16 [n*g*g, D]
 * (class-token rows dropped). */
int pcg_embed_bwd(const float *dx0, const void *dx0_bf16, const float *v, const float *gamma, void *d_patch_bf16,
                  int n, int T, int D, void *stream);

/* ---- short-sequence multi-head attention, head dim 64, no mask ------------------------------------------
 * replaces nn.MultiheadAttention(need_weights=False) core (ruclip/model.py:43-49).
 * qkv bf16 [n*T, 3D] (q pre-scaled), out bf16 [n*T, D], lse f32 [n, heads, T] (natural log). */
int pcg_attn_fwd(const void *qkv, void *out, float *lse, int n, int T, int heads, void *stream);
/* delta_ws: pcg_attn_bwd_workspace_bytes(n, T, heads) bytes of scratch (rowsum(dO*O) per query; for T > 257 also the
 * fp32 dQ sums that the key-tile CTAs of the tcgen05 backward reduce into). */
size_t pcg_attn_bwd_workspace_bytes(int n, int T, int heads);
int pcg_attn_bwd(const void *qkv, const void *out, const void *d_out, const float *lse, float *delta_ws,
                 void *d_qkv, int n, int T, int heads, void *stream);
/* Wide heads (head dim 80 / 88 -- ViT-H/14, the default of perceptor/losses/open_clip.py:8-12, and ViT-g/14 --
 * padded to 128 columns per head): qkv bf16 [n*T, 3*heads*128], out / d_out bf16 [n*T, heads*128], q pre-scaled by
 * 1/sqrt(head dim), pad columns zero.  Same lse / delta_ws conventions as above. */
int pcg_attn_fwd_wide(const void *qkv, void *out, float *lse, int n, int T, int heads, void *stream);
int pcg_attn_bwd_wide(const void *qkv, const void *out, const void *d_out, const float *lse, float *delta_ws,
                      void *d_qkv, int n, int T, int heads, void *stream);

/* Class-token attention of the LAST block: query row 0 of every (cutout, head) against all T keys
 * (ruclip/model.py:43-49 restricted to the one row that ruclip/model.py:126 reads).  head_stride = columns per head
 * in qkv / out (pcg_head_stride).  fwd writes out[n*T + 0, :] and lse[n, :, 0] only; bwd writes ALL of d_qkv
 * (dK, dV of every row, dQ of row 0, zeros in the other dQ rows) from d_out[n*T + 0, :]. */
int pcg_attn_cls_fwd(const void *qkv, void *out, float *lse, int n, int T, int heads, int head_stride, void *stream);
int pcg_attn_cls_bwd(const void *qkv, const void *out, const void *d_out, const float *lse, void *d_qkv, int n, int T,
                     int heads, int head_stride, void *stream);

/* ---- head: ln_post(CLS) @ proj -> L2 normalise -> spherical distance loss, forward AND gradient ---------
 * replaces ruclip/model.py:126-129 + F.normalize (models/open_clip.py:120-121) + CLIP.forward
 * (losses/clip/clip.py:89-99).  x f32 [n*T, D]; targets f32 [M,E]; tweights f32 [M].
 * loss_sum: one f32, atomically += sum over this call's cutouts of sum_m w_m d(n,m) * scale.
 * enc_out (nullable) f32 [n,E] normalised (or raw if !normalize) encodings.
 * d_enc (nullable) f32 [n,E]: an upstream gradient w.r.t. the encodings; when given it replaces the loss
 *   gradient (this is autograd through encode_images for callers other than the CLIP loss).
 * dx (nullable) f32 [n*T, D]: CLS rows receive d(loss)/dx, all other rows are zeroed. dx_bf16 (nullable) likewise;
 *   either one (or both) requests the gradient.
 * workspace: pcg_head_workspace_bytes(n, D, E) bytes of device scratch (the batched projection's operands). */
size_t pcg_head_workspace_bytes(int n, int D, int E);
int pcg_head_loss(const float *x, const float *ln_g, const float *ln_b, const float *proj, const float *targets,
                  const float *tweights, int n, int T, int D, int E, int M, float scale, int normalize,
                  float *loss_sum, float *enc_out, const float *d_enc, float *dx, void *dx_bf16, float *workspace,
                  void *stream);

/* ---- guided-sampling glue ----------------------------------------------------------------
 * out[n,i] = a[n] x[n,i] + b[n] clamp(y[n,i], -clamp_limit, clamp_limit) + c[n];  coef = f32 [3][n_samples] (a,b,c)
 * on the device, y nullable (then b is ignored), clamp_limit = INFINITY for no clamp.  With coefficients from
 * alpha = cos(t pi/2), sigma = sin(t pi/2) this is Predictions.denoised_images / .step(eta=0) / .guided and their
 * backward (perceptor/models/velocity_diffusion/predictions.py:50-66,68-105,148-155; diffusion_space.py:1-6). */
int pcg_affine2(const float *x, const float *y, const float *coef, float *out, int n_samples, long long per,
                float clamp_limit, void *stream);
/* g == NULL: out = clamp(x, lo, hi).  g != NULL: out = g * [g (x - clamp(x, lo, hi)) >= 0], the backward of
 * clamp_with_grad (perceptor/transforms/clamp_with_grad.py:8-27). */
int pcg_clamp_with_grad(const float *x, const float *g, float *out, long long total, float lo, float hi,
                        void *stream);

/* ---- whole path ------------------------------------------------------------------------------------------ */
/* bytes of scratch (transient) and stash (activations kept for backward) for n cutouts. */
size_t pcg_workspace_bytes(const pcg_vit_config *cfg, int n_cut);
size_t pcg_stash_bytes(const pcg_vit_config *cfg, int n_cut);

typedef struct pcg_guidance_args {
    const pcg_vit_config *cfg;
    const pcg_vit_weights *w;
    const float *images; int32_t B, H, W;
    const int32_t *cuts; int32_t n_cut; int32_t max_in_w;
    const pcg_resize_tables *tabs;
    const float *mean_host, *std_host;
    const float *targets, *tweights; int32_t n_targets;
    float loss_scale;        /* multiplier / (N_total * M)                              */
    void *workspace; size_t workspace_bytes;
    void *stash; size_t stash_bytes;      /* may be NULL when want_grad == 0             */
    int32_t want_grad;       /* 1: keep activations for pcg_guidance_bwd                */
    float *loss_sum;         /* one f32, caller zeroes                                  */
    float *enc_out;          /* nullable f32 [n_cut,E]                                  */
    int32_t normalize;       /* L2-normalise encodings (encode_images(normalize=...))   */
    float *d_images;         /* bwd only: f32 [B,3,H,W], caller zeroes                  */
    const float *d_enc;      /* bwd only, nullable: f32 [n_cut,E] upstream gradient of the encodings
                                (replaces the loss gradient; targets may then be NULL)   */
} pcg_guidance_args;

/* forward: sampler -> patch embed -> L x block -> head (+ loss).  replaces CLIP.forward end to end. */
int pcg_guidance_fwd(const pcg_guidance_args *a, void *stream);
/* backward to the image: replaces autograd through everything above (dgrad only; weights are frozen,
 * perceptor/models/open_clip.py:73-76). */
int pcg_guidance_bwd(const pcg_guidance_args *a, void *stream);
/* The loss reads only the class-token row of the last block's output (ruclip/model.py:126), so by default
 * pcg_guidance_fwd / _bwd run the last block's attention output, out-projection, ln_2 and MLP on the n class-token rows
 * instead of all n*T rows (K and V projections, ln_1 and everything below stay full): the same loss and image gradient
 * from 10/12 of one block's GEMM work less.  on = 0 restores the full last block (also: PCG_FULL_LAST_BLOCK=1 in the
 * environment); returns the previous setting.  Takes effect for subsequent calls (captured CUDA graphs keep theirs).
 * A pcg_guidance_bwd must run under the setting its pcg_guidance_fwd ran under: the forward keeps only the rows the
 * backward of the same setting reads. */
int pcg_set_pooled_last_block(int on);
/* number of kernels the last fwd / bwd call launched (for bench.py's gpu_launches). */
int pcg_last_launch_count(void);

/* ---- diagnostics: device-side timing per kernel family (CUDA events on the launch stream) -------------- */
enum {
    PCG_PROF_GEMM = 0, PCG_PROF_ATTN_FWD = 1, PCG_PROF_ATTN_BWD = 2, PCG_PROF_LAYERNORM = 3, PCG_PROF_SAMPLER_FWD = 4,
    PCG_PROF_SAMPLER_BWD = 5, PCG_PROF_EMBED = 6, PCG_PROF_HEAD = 7, PCG_PROF_KINDS = 8
};
int pcg_profile_enable(int on);
/* ms/work/count: arrays of PCG_PROF_KINDS; work = algorithmic FLOPs (GEMM, attention) or bytes (the rest). */
int pcg_profile_collect(double *ms, double *work, int *count);

#ifdef __cplusplus
}
#endif
#endif /* PCG_H_ */
