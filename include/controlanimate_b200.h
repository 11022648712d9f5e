/*
 * controlanimate_b200 — C ABI of the B200 (sm_100a) denoising hot path.
 *
 * Drop-in boundary for the three operator interfaces of intellerce/controlanimate's denoising
 * loop.  The reference is pure Python/PyTorch: there is no FFI in it, so each entry
 * point cites the reference *operator* it replaces (paths relative to the reference repo root).
 * Host code (the Python modules under controlanimate_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns ca_status_t (0 = ok) and never throws; ca_last_error() gives the text
 *   - all pointers are DEVICE pointers unless the parameter is documented as "host"
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no internal synchronisation,
 *     no device allocation: scratch space is provided by the caller (see *_workspace_bytes)
 *   - dtype is the storage type of activations; accumulation is always fp32 (or wider)
 *   - layouts of a video activation with logical shape [b, c, f, h, w]:
 *       CA_LAYOUT_NCFHW  memory order b,c,f,h,w  (what the reference passes between modules)
 *       CA_LAYOUT_BFHWC  memory order b,f,h,w,c  (native: "(b f) h w c" = token-major rows of c)
 */
#ifndef CONTROLANIMATE_B200_H_
#define CONTROLANIMATE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { CA_OK = 0, CA_ERR_INVALID = 1, CA_ERR_UNSUPPORTED = 2, CA_ERR_CUDA = 3 } ca_status_t;
typedef enum { CA_BF16 = 0, CA_F16 = 1, CA_F32 = 2 } ca_dtype_t;
typedef enum { CA_LAYOUT_NCFHW = 0, CA_LAYOUT_BFHWC = 1 } ca_layout_t;

#define CA_MAX_NETS 8      /* ControlNets per Multi-ControlNet set */
#define CA_MAX_RESIDUALS 16 /* 12 down + 1 mid for SD1.5 */

/* Library / device introspection. */
const char* ca_version(void);
const char* ca_last_error(void);
/* 10*major+minor of the current device (100 on B200); negative on error. */
int ca_device_sm(void);

/* ---------------------------------------------------------------------------------------------
 * Kernel (2): y = SiLU(GroupNorm(x + temb[b, c])).
 * Replaces InflatedGroupNorm.forward + F.silu (animatediff/models/resnet.py:23-31, 191-192,
 * 199-208) and conv_norm_out + conv_act (animatediff/models/unet.py:614-615); with apply_silu = 0
 * also the transformer-entry GroupNorms (motion_module.py:144, attention.py:131).
 *   x, y      [b,c,f,h,w] in `layout`, dtype `dtype`; y may alias x
 *   gamma,beta[c] fp32;  temb [b,c] fp32 with row stride temb_ld (elements; 0 = c) or NULL (the projected time
 *             embedding, resnet.py:196-200; the caller may fold conv1's bias into it)
 *   per_frame 1: statistics per (b, f, group) over (c/groups, h, w)  (use_inflated_groupnorm, v2)
 *             0: statistics per (b, group) over (c/groups, f, h, w)  (plain nn.GroupNorm, v1)
 *   workspace ca_groupnorm_workspace_bytes(...) bytes of scratch (may be NULL if that is 0)
 * ------------------------------------------------------------------------------------------- */
size_t ca_groupnorm_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int layout,
                                    int dtype);
int ca_groupnorm_silu(const void* x, void* y, const float* gamma, const float* beta, const float* temb,
                      long long temb_ld, int b, int c, int f, int h, int w, int groups, float eps, int per_frame,
                      int apply_silu, int layout, int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel (3): single-pass Multi-ControlNet residual merge.
 *   dst_i  (=|+=)  sum_k scales[k*n_res + i] * res[k*n_res + i]        for i in [0, n_res)
 * Replaces the per-net conditioning-scale multiply and the sum over nets (diffusers 0.23.0
 * ControlNetModel/MultiControlNetModel.forward, called at modules/controlresiduals_pipeline.py:
 * 294-302), the 13 '(b f) c h w -> b c f h w' rearranges (:304-312) and, with add_into_dst = 1,
 * the skip additions animatediff/models/unet.py:567-576, 584-585.
 *   res     host array [n_nets*n_res] of device pointers, each a residual of shape
 *           [(b_res f), c_i, h_i, w_i]: NCHW per frame when layout == CA_LAYOUT_NCFHW (what
 *           diffusers ControlNets emit), or [(b_res f), h_i, w_i, c_i] when CA_LAYOUT_BFHWC
 *   scales  host array [n_nets*n_res] fp32 (guess-mode logspace factors folded in by the caller)
 *   dst     host array [n_res] of device pointers: [b_dst, c_i, f, h_i, w_i] in `layout`
 *   chw     host array [n_res*3] = c_i, h_i, w_i
 *   b_res   1 or b_dst (batch broadcast, unet.py:572 in guess mode + CFG)
 *   add_into_dst  0: dst = sum (contract-preserving producer);  1: dst += sum (in-place on skips)
 * ------------------------------------------------------------------------------------------- */
int ca_residual_merge(const void* const* res, const float* scales, void* const* dst, const int* chw, int n_nets,
                      int n_res, int b_res, int b_dst, int f, int add_into_dst, int layout, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel (1) pieces, all on token-major activations: row t = (b*f + frame)*d + site, d = h*w.
 * ------------------------------------------------------------------------------------------- */

/* y[t,:] = LayerNorm(x[t,:]) * gamma + beta (+ pe[frame(t), :]).
 * Replaces nn.LayerNorm (motion_module.py:214, 221) and, with pe != NULL, the positional-encoding
 * add of VersatileAttention.forward (motion_module.py:285-288, PositionalEncoding :227-245), done
 * in token order so the '(b f) d c -> (b d) f c' rearrange never materialises.
 *   x,y [rows, c] dtype; gamma,beta [c] fp32; pe [>=f, c] fp32 or NULL; rows = b*f*d  */
int ca_layernorm_pe(const void* x, void* y, const float* gamma, const float* beta, const float* pe, long long rows,
                    int c, int f, int d, float eps, int dtype, void* stream);

/* Temporal self-attention core: for every (b, site, head): O = softmax(Q K^T * scale) V over the
 * f frames.  Replaces the head split + attention + head merge of the AttentionProcessor
 * (modules/attention_processor.py:56-62 / :247-256; xformers memory_efficient_attention on the
 * reference's default GPU path) for VersatileAttention's temporal mode (motion_module.py:285,321,327).
 *   q,k,v  token-major [b*f*d, *] with row strides ldq/ldk/ldv (elements); head hh occupies
 *          columns [hh*head_dim, (hh+1)*head_dim).  A packed [T, 3C] QKV buffer is passed as
 *          q = base, k = base + C, v = base + 2C, ld* = 3C.
 *   o      [b*f*d, heads*head_dim] with row stride ldo
 *   seq_major 0: rows are token-major, t = (b*f + frame)*d + site (native layout)
 *             1: rows are the reference's "(b d) f c" order, t = (b*d + site)*f + frame — exactly the
 *                hidden_states an AttentionProcessor receives from VersatileAttention (boundary B1)
 *   constraints: f <= 32, head_dim % 8 == 0, head_dim <= 256, dtype bf16/f16, 16-byte aligned rows */
int ca_temporal_attn_core(const void* q, const void* k, const void* v, void* o, int b, int f, int d, int heads,
                          int head_dim, long long ldq, long long ldk, long long ldv, long long ldo, int seq_major,
                          float scale, int dtype, void* stream);

/* Kernel (1), fused: one whole temporal-attention block in ONE launch,
 *   y = x + to_out( attention( LayerNorm(x) * gamma + beta + pe[frame] ) ) + b_out
 * Replaces nn.LayerNorm (motion_module.py:214), VersatileAttention.forward (:272-329: both rearranges, the positional
 * encoding :287-288, the processor call :321), the AttentionProcessor arithmetic (modules/attention_processor.py:186-272:
 * to_q/to_k/to_v, softmax(q k^T scale) v per head over the f frames of every site, to_out[0]) and the residual add of
 * TemporalTransformerBlock.forward (:219).  x is read once and y written once; QKV and out projections run on tcgen05.
 *   x, y       token-major [b*f*d, C] dense rows (row t = (b*f + frame)*d + site); y may alias x
 *   ln_gamma, ln_beta [C] fp32; pe [>= f, C] fp32 or NULL; bo [C] fp32
 *   wqkv_perm  [heads * nq, C] (dtype), nq = 3*head_dim rounded up to a multiple of 16: per head the rows of to_q, to_k,
 *              to_v of that head followed by zero rows (the host packs it once per weight version)
 *   wo         [C, C] (dtype) = to_out[0].weight
 *   constraints: C in {64, 128, 320} (returns CA_ERR_UNSUPPORTED otherwise: 128 rows of the normalised tile and of the
 *   attention output must both fit one SM's shared memory), head_dim % 8 == 0, f <= 32, dtype bf16/f16 */
int ca_temporal_attn_fused(const void* x, void* y, const float* ln_gamma, const float* ln_beta, const float* pe,
                           const void* wqkv_perm, const void* wo, const float* bo, int b, int f, int d, int C, int heads,
                           float eps, float scale, int dtype, void* stream);

/* Cross-attention core of the spatial transformer: every latent site attends to the
 * kv_len <= 96 prompt tokens, O = softmax(Q K^T * scale) V per (frame, head).
 * Replaces the attention arithmetic of BasicTransformerBlock.attn2 (reference animatediff/models/attention.py:283-289
 * -> modules/attention_processor.py:56-62 baddbmm/softmax/bmm, :247-256 SDPA).
 *   q, o   [n_frames * d, heads*head_dim] token-major rows (row strides ldq / ldo elements), token = frame*d + site
 *   k, v   [n_ctx, kv_len, heads*head_dim] rows (row strides ldk / ldv, prompt strides ctx_stride_k / ctx_stride_v);
 *          they may be column slices of one fused [n_ctx, kv_len, 2C] projection
 *   ctx_of_frame  device int32 [n_frames] prompt index of every frame, or NULL: frame n uses prompt n / (n_frames/n_ctx)
 *   constraints: head_dim in {40, 80, 160} (the SD1.5 widths), kv_len <= 96, dtype bf16/f16, strides multiples of 8
 *   returns CA_ERR_UNSUPPORTED for other head_dim / kv_len (the caller keeps its library path for those) */
int ca_cross_attn_core(const void* q, const void* k, const void* v, void* o, int n_frames, int d, int heads, int head_dim,
                       int n_ctx, int kv_len, long long ldq, long long ldk, long long ldv, long long ldo,
                       long long ctx_stride_k, long long ctx_stride_v, const int* ctx_of_frame, float scale, int dtype,
                       void* stream);

/* Dense projection on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA operands):
 *   y = epilogue(x @ w^T + bias) (+ residual)
 * Replaces the nn.Linear projections of the motion module: to_q/to_k/to_v (fused as one [3C, C]
 * weight), to_out[0] (+bias, +residual: motion_module.py:215-219), proj_in/proj_out (:147, :155),
 * GEGLU ff.net.0.proj and ff.net.2 (:221).
 *   x [m, k] row stride ldx; w [n, k] row-major (nn.Linear layout); bias [n] fp32 or NULL;
 *   residual [m, n_out] row stride ldr or NULL; y [m, n_out] row stride ldy
 *   epilogue: CA_EPI_NONE n_out = n;  CA_EPI_GEGLU n_out = n/2: y = a * gelu_erf(g) where a/g are
 *   columns j and j + n/2 of the product (diffusers GEGLU chunk(2, -1))
 *   constraints: k % 8 == 0, n % 32 == 0 (n % 64 == 0 for GEGLU), 16-byte aligned rows, dtype bf16/f16 */
typedef enum { CA_EPI_NONE = 0, CA_EPI_GEGLU = 1 } ca_epilogue_t;
int ca_linear(const void* x, const void* w, const float* bias, const void* residual, void* y, long long m, int n,
              int k, long long ldx, long long ldr, long long ldy, int epilogue, int dtype, void* stream);

/* LayerNorm folded into the projection that consumes it.  With w_gain = w * diag(gamma) (rounded to dtype),
 * colsum[n] = sum_k w_gain[n][k] and shift[r][n] = sum_k w[n][k] * (beta[k] + pe[r][k]) (+ bias[n]):
 *   Linear(LayerNorm(x) + pe[frame])[row, n] = rstd[row] * (x[row] . w_gain[n] - mean[row] * colsum[n]) + shift[frame(row)][n]
 * so the normalised tensor never exists in HBM: ca_row_stats reads x once and writes 8 bytes per row, ca_linear_ln runs the
 * tcgen05 GEMM on the RAW rows and applies (mean, rstd) in its epilogue (also in front of GEGLU).
 * Replaces nn.LayerNorm + PositionalEncoding + to_q/to_k/to_v (motion_module.py:214-215, 285-288, 321) and
 * norm1/norm2/norm3 + the projections behind them (animatediff/models/attention.py:271-297).
 *   stats [rows] (mean, rstd) fp32 pairs; shift [shift_rows, n] fp32 with shift_rows == 1 or >= frames, row
 *   (token_row / sites) % frames is used (token rows are (b f d)-major; sites % 32 == 0 when shift_rows > 1) */
int ca_row_stats(const void* x, float* stats, long long rows, int c, long long ldx, float eps, int dtype, void* stream);
int ca_linear_ln(const void* x, const void* w_gain, const float* colsum, const float* shift, int shift_rows, int frames,
                 int sites, const float* stats, void* y, long long m, int n, int k, long long ldx, long long ldy,
                 int epilogue, int dtype, void* stream);

/* Convolution epilogue on channels-last rows:  y = (act(x + bias) + residual) * scale.
 * Replaces the broadcast bias add behind every InflatedConv3d / nn.Conv2d of the path
 * (animatediff/models/resnet.py:12-20), the shortcut add and 1/output_scale_factor of
 * ResnetBlock3D.forward (resnet.py:213-216) and conv+SiLU of the ControlNet conditioning embedding
 * (diffusers 0.23.0 ControlNetConditioningEmbedding, reached from modules/controlresiduals_pipeline.py:294-302).
 *   x, y, residual [rows, c] dense rows, dtype; y may alias x or residual; bias [c] fp32 or NULL;
 *   residual NULL or same shape as x; act 0 = none, 1 = SiLU; c % (16 / sizeof(dtype)) == 0 */
int ca_bias_act_residual(const void* x, const float* bias, const void* residual, void* y, long long rows, int c,
                         float scale, int act, int dtype, void* stream);

/* Spatial self-attention core: o = softmax(q k^T * scale) v over the `sites` (h*w) tokens of one frame, per head — the
 * arithmetic of BasicTransformerBlock.attn1 (animatediff/models/attention.py:268-271) behind the reference's AttentionProcessor
 * (modules/attention_processor.py:56-62 / :247-256).  tcgen05 / TMEM flash attention: scores and the output accumulator live in
 * TMEM, one thread per query row computes the exponentials, P V consumes V as it lies in HBM (MN-major operand).
 *   q, k, v: rows (frame * sites + site) with row strides ldq / ldk / ldv (elements), head h at columns [h * head_dim, ...)
 *   — e.g. the three column blocks of the packed [T, 3C] projection output; o likewise with ldo.  head_dim % 8 == 0, <= 64. */
int ca_spatial_attn_core(const void* q, const void* k, const void* v, void* o, int frames, int sites, int heads, int head_dim,
                         long long ldq, long long ldk, long long ldv, long long ldo, float scale, int dtype, void* stream);

/* The two pure data-movement steps of the UNet's up path, on channels-last rows ([n, h, w, c] memory order):
 *   ca_upsample_nearest: F.interpolate(mode="nearest") of Upsample3D.forward (animatediff/models/resnet.py:63-69);
 *     exact_2x = 1 -> scale_factor 2 (out = 2 * in), else to the explicit (out_h, out_w) of forward_upsample_size
 *     (unet.py:491-499, 596-597) with torch's source index min(floor(dst * in / out), in - 1)
 *   ca_concat_channels: torch.cat([hidden_states, res_hidden_states], dim=1) in front of every up-block resnet
 *     (animatediff/models/unet_blocks.py:636, :742): y[r] = [a[r], b[r]] for rows = n * h * w
 *   c, ca, cb multiples of 16 / sizeof(dtype); 16-byte aligned pointers; x / a / b / y dense */
int ca_upsample_nearest(const void* x, void* y, long long n, int c, int in_h, int in_w, int out_h, int out_w, int exact_2x,
                        int dtype, void* stream);
int ca_concat_channels(const void* a, const void* b, void* y, long long rows, int ca, int cb, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loop glue: classifier-free-guidance combine + DDIM update of one step in ONE launch.
 * Replaces `noise_pred.to(latents_dtype)`, `noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)`
 * (animatediff/pipelines/controlanimation_pipeline.py:841, 845-846) and `scheduler.step(...).prev_sample` (:849; diffusers
 * 0.23.0 DDIMScheduler.step with eta = 0, epsilon prediction, clip_sample False — third party, restated):
 *     eps = cfg ? u + guidance * (c - u) : model_out;  x0 = (x - sqrt(1 - a_t) eps) / sqrt(a_t);
 *     latents_out = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps
 *   model_out   [cfg ? 2n : n] elements of `model_dtype`, dense: the UNet output rows [uncond | cond]
 *   latents     [n] elements of `latent_dtype`; latents_out [n] (may alias latents); noise_out [n] or NULL (the guided eps)
 *   fp32 arithmetic; model_out is first rounded to latent_dtype as the reference's `.to(latents_dtype)` does
 * ------------------------------------------------------------------------------------------- */
int ca_cfg_ddim_step(const void* model_out, const void* latents, void* latents_out, void* noise_out, long long n, int cfg,
                     float guidance, float sqrt_alpha_t, float sqrt_one_minus_alpha_t, float sqrt_alpha_prev,
                     float sqrt_one_minus_alpha_prev, int model_dtype, int latent_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTROLANIMATE_B200_H_ */
