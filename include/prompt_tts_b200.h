/*
 * prompt_tts_b200 -- C ABI of the B200 (sm_100a) kernels behind the prompt-tts denoiser hot path.
 *
 * The reference (khaidoan25/prompt-tts) has no FFI layer: its hot path is Python nn.Modules over
 * PyTorch library kernels.  This header is the boundary a maintainer would bind instead of those
 * library calls (ctypes stub in INTEGRATION.md).  Every entry point
 *   - is plain `extern "C"`: raw device pointers, sizes, a cudaStream_t passed as void*;
 *   - enqueues on the caller's stream, never synchronises, never allocates device memory;
 *   - returns 0 on success, a negative PT_E* code otherwise (message via pt_last_error()).
 * Activations are channels-last: [B, L, C] (C contiguous), bf16 unless stated.
 * Parameters stay in the reference layout (fp32); packed bf16 copies are made by pt_pack_*.
 *
 * Reference interface replaced, per group (file:line into the reference tree):
 *   pt_gemm            Conv1d k3/k1 + Linear + SDPA contractions   tts/ldm/resnet.py:171,193,226-228,253,276,279
 *                                                                  tts/ldm/transformer_1d.py:134,253,258-265
 *                                                                  (diffusers Attention/FeedForward, unvendored)
 *   pt_groupnorm_*     GroupNorm(+SiLU)                            tts/ldm/resnet.py:169,189,238-240,267-273
 *                                                                  tts/ldm/transformer_1d.py:130,251; unet_1d_condition.py:401,732-733
 *   pt_layernorm_*     LayerNorm in BasicTransformerBlock          (diffusers 0.15, used at transformer_1d.py:258-265)
 *   pt_softmax_*       softmax(QK^T/sqrt d)                        (diffusers AttnProcessor2_0)
 *   pt_attn_fwd/bwd    fused softmax(QK^T/sqrt d) V                (diffusers AttnProcessor2_0 / F.scaled_dot_product_attention)
 *   pt_geglu_*         h * gelu_erf(g)                             (diffusers GEGLU)
 *   pt_conv_in_*, pt_conv_out_*   8<->C convs                      tts/ldm/unet_1d_condition.py:193,410,654,734
 *   pt_upsample2_*     nearest x2                                  tts/ldm/resnet.py:41-44
 *   pt_time_sinusoid   Timesteps                                   tts/ldm/unet_1d_condition.py:209,622
 *   pt_text_embed_*    Embedding + transposed PE                   tts/models.py:32-52,112-115
 *   pt_rvq_encode_ws   encodec ResidualVectorQuantization.encode   data_preparation/generate_code.py:48
 *   pt_rvq_decode      encodec ResidualVectorQuantization.decode   decode_codec.py:16
 *   pt_codes_affine    codes/1023 -> Normalize(0.5,0.5)            tts/dataloader.py:64,77,168-170
 *   pt_add_noise, pt_mse_*        train-step glue                  train.py:96-98,107
 *   pt_ddpm_step       sampling step (diffusers DDPMScheduler.step; SURVEY 8 row N1, not in the reference tree)
 */
#ifndef PROMPT_TTS_B200_H
#define PROMPT_TTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_OK 0
#define PT_EINVAL (-1)   /* bad shape / dtype / alignment */
#define PT_ECUDA (-2)    /* CUDA runtime or driver error */
#define PT_EARCH (-3)    /* not an sm_100 device */

int pt_version(void);
const char* pt_last_error(void);
/* 0 if device `dev` is sm_100 and the TMA driver entry point resolves */
int pt_check_device(int dev);
/* leave n SMs free of persistent (GEMM / attention) CTAs so that an overlapped collective can make progress (data-parallel runs) */
int pt_set_sm_reserve(int n);
/* cumulative number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long pt_launch_count(void);
/* tile configuration the last pt_gemm call of this thread chose: block_n | (cluster of two ? 0 : 1) -- the value that, passed as
   pt_gemm_t.block_n, reproduces it (tools/gemm_sweep.py: is a forced width a different configuration from the library's choice?) */
int pt_gemm_last_tile(void);

/* ------------------------------------------------------------------ tcgen05 GEMM family ------- */
/* One operand = a bf16 tensor of rank <= 4.  dim[0] is the contiguous axis (stride[0] == 1).
 * kmajor=1: dim[0] is the contraction axis K, dim[1] the M (or N) axis.
 * kmajor=0: dim[0] is the M (or N) axis, dim[1] the contraction axis K.
 * dim[2], dim[3]: batch axes.  Strides in ELEMENTS, multiples of 8 (16 bytes).
 * Reads outside [0, dim) return zero (TMA out-of-bounds fill) -- this is how conv padding,
 * ragged tiles and K tails are handled. */
typedef struct {
  const void* ptr;
  int64_t dim[4];
  int64_t stride[4];
  int32_t kmajor;
  int32_t batched; /* 1: dim[2],dim[3] indexed by the launch's (z2,z3); 0: coordinate 0 */
} pt_operand_t;

/* The contraction is a list of segments; every segment contributes
 *   sum_{rep < nrep} sum_{k < nk}  A[m + a_mn_shift, a_k0 + k ; c2 = rep_c2_0 + rep] * B[n + b_mn_shift, b_k0 + k ; same c2]
 * (c2 from `rep` only when nrep_is_batch=1, else from the launch's z2).
 * conv k3 forward = 3 segments (taps) with a_mn_shift = -1,0,+1; weight-gradient = 1 segment with
 * nrep = batch; skip-concat / fused shortcut = extra segments on a second A map. */
typedef struct {
  int32_t a_idx, b_idx;   /* which of a[2] / b[2] */
  int32_t a_k0, b_k0;     /* may be negative (reads zero) */
  int32_t a_mn_shift, b_mn_shift;
  int32_t nk;             /* contraction length of this segment (elements) */
  int32_t nrep;           /* >= 1 */
  int32_t rep_is_batch;   /* 1: c2 := rep_c2_0 + rep (reduction over dim[2]) */
  int32_t rep_c2_0;
  int32_t b_k0_z2;        /* b_k0 += z2 * b_k0_z2: one launch covers the three taps of a conv weight gradient (z2 = tap) */
} pt_segment_t;

#define PT_OUT_BF16 0
#define PT_OUT_F32 1
#define PT_OUT_F32_ATOMIC_ADD 2

typedef struct {
  pt_operand_t a[2];
  pt_operand_t b[2];
  pt_segment_t seg[8];
  int32_t nseg;
  int32_t M, N;            /* output tile space per (z2,z3) */
  int32_t nz2, nz3;        /* batch extents of the output tile space */
  int32_t splitk;          /* ignored (kept for ABI stability): PT_OUT_F32_ATOMIC_ADD outputs are scheduled stream-K by the library */
  int32_t block_n;         /* 0 = auto; else 64 / 96 (K-major B only) / 128 / 160 / 192 / 224 / 256; bit 0 set (e.g. 257) = never pair CTAs into multicast clusters */
  /* epilogue:  out = alpha * acc + bias[n] + bias_z2[z2, n] + residual[z2,z3,m,n] */
  void* out;
  int32_t out_dtype;
  int64_t out_stride_m, out_stride_z2, out_stride_z3; /* elements.  PT_OUT_BF16: `out` 16-byte aligned, N and the strides multiples
                                                        * of 8 -- the tiles are written by TMA stores through a rank-4 tensor map
                                                        * (N, M, nz2, nz3); `residual` may alias `out` (in-place accumulation) */
  float alpha;
  const float* bias;       /* [N] fp32 or NULL */
  const float* bias_z2;    /* [nz2, >=N] fp32 rows of stride bias_z2_stride, or NULL (time-embedding shift) */
  const void* residual;    /* bf16, same indexing as out with its own strides, or NULL */
  int64_t res_stride_m, res_stride_z2, res_stride_z3;
  int64_t bias_z2_stride;  /* elements; 0 means N */
  int64_t out_stride_n;    /* PT_OUT_F32_ATOMIC_ADD only: element stride between output columns; 0 means 1 */
  int32_t out_transposed;  /* PT_OUT_BF16 only.  1: `out` and `residual` are indexed [z2, z3, n, m] -- the M index is the contiguous one and
                            * out_stride_m / res_stride_m are the element strides of the N index -- and bias / bias_z2 are indexed by m.
                            * This is how a convolution over few rows per sample runs with the WEIGHTS on the 128-row side of the tile:
                            * M = C_out (a multiple of 128), N = L (tile width = L rounded up to 16: block_n 96 / 192 for 94 / 188 rows)
                            * instead of 128-row tiles that are 27 % padding.  M and the strides must be multiples of 8. */
} pt_gemm_t;

int pt_gemm(const pt_gemm_t* g, void* stream);

/* ------------------------------------------------------------------ normalisation ------------- */
/* GroupNorm over channels-last x[B, L, C] bf16: stats[B, G, 2] = (mean, rstd) fp32. */
int pt_groupnorm_stats(const void* x, float* stats, int B, int L, int C, int G, float eps, void* stream);
/* y = act(gn(x) * gamma + beta); act: 0 none, 1 SiLU */
int pt_groupnorm_apply(const void* x, const float* stats, const float* gamma, const float* beta, void* y,
                       int B, int L, int C, int G, int act, void* stream);
/* stats + apply in one call (tts/ldm/resnet.py:238-240, 267-273: GroupNorm -> SiLU).  When one sample [L, C] fits in the shared
 * memory of a cluster of 8 CTAs (<= 100 KB each) it is read from HBM once (bulk copy -> reduce -> exchange of the group sums over distributed shared
 * memory -> normalise from the resident copy); otherwise the two calls above.  stats[B, G, 2] is written for the backward pass. */
int pt_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats,
                     int B, int L, int C, int G, float eps, int act, void* stream);
/* backward of apply+stats: dx bf16 (+ dx_add if not NULL: fused accumulation of the gradient that reached x through another
 * branch; dx may alias dx_add); dgamma/dbeta fp32 [C] are ACCUMULATED (atomic add).  scratch: fp32 [B, G, 2], zero-filled by the callee. */
int pt_groupnorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta,
                     const void* dx_add, void* dx, float* dgamma, float* dbeta, float* scratch,
                     int B, int L, int C, int G, int act, void* stream);

/* LayerNorm over rows x[M, C] bf16, eps; rowstats[M,2] = (mean, rstd) */
int pt_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* rowstats,
                     int64_t M, int C, float eps, void* stream);
/* dx = LN backward (+ dx_add if not NULL: fused residual-gradient add); dgamma/dbeta accumulated */
int pt_layernorm_bwd(const void* dy, const void* x, const float* rowstats, const float* gamma,
                     const void* dx_add, void* dx, float* dgamma, float* dbeta,
                     int64_t M, int C, void* stream);

/* ------------------------------------------------------------------ attention pieces ---------- */
/* P[r, :n] = softmax(S[r, :n]) ; S fp32 rows of stride ld_s, P bf16 rows of stride ld_p */
int pt_softmax_fwd(const float* S, void* P, int64_t rows, int n, int64_t ld_s, int64_t ld_p, void* stream);
/* dS[r, :] = scale * P * (dP - sum(dP * P)) ; dP fp32 (stride ld_s), P bf16, dS bf16 (stride ld_p) */
int pt_softmax_bwd(const float* dP, const void* P, void* dS, int64_t rows, int n, int64_t ld_s, int64_t ld_p,
                   float scale, void* stream);
/* Fused softmax attention on tcgen05 (no mask, no dropout: diffusers AttnProcessor2_0 as used from
 * tts/ldm/transformer_1d.py:258-265 and tts/models.py:95-100).  q/k/v/o/d_o/dq/dk/dv point at the head-0 column of
 * bf16 [B, L, W] tensors (row stride *_rs, batch stride *_bs, in elements; head h occupies columns [h*d, (h+1)*d)),
 * so fused QKV / KV projections are consumed and their gradients produced in place.  d: multiple of 8, <= 192.
 * lse [B, H, Lq] fp32 = log-sum-exp of the scaled logits (written by fwd, read by bwd); delta: bwd scratch [B, H, Lq].
 * pt_attn_bwd: Lq <= 1984 (the per-query statistics of one (batch, head) are staged in shared memory). */
typedef struct {
  const void* q; int64_t q_rs, q_bs;
  const void* k; const void* v; int64_t kv_rs, kv_bs;
  void* o; int64_t o_rs, o_bs;
  float* lse;
  const void* d_o; int64_t do_rs, do_bs;
  void* dq; int64_t dq_rs, dq_bs;
  void* dk; void* dv; int64_t dkv_rs, dkv_bs;
  float* delta;
  int32_t B, H, Lq, Lk, d;
  float scale;
} pt_attn_t;
int pt_attn_fwd(const pt_attn_t* a, void* stream);
int pt_attn_bwd(const pt_attn_t* a, void* stream);
/* GEGLU: y[m, j] = u[m, j] * gelu_erf(u[m, F + j]), u[M, 2F] bf16 -> y[M, F] bf16 */
int pt_geglu_fwd(const void* u, void* y, int64_t M, int F, void* stream);
int pt_geglu_bwd(const void* dy, const void* u, void* du, int64_t M, int F, void* stream);

/* ------------------------------------------------------------------ elementwise / layout ------ */
int pt_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream);
int pt_silu_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);             /* y = bf16(silu(x)) */
int pt_silu_bwd_f32(const float* x, const float* dy, float* dx, int64_t n, void* stream);
/* strided 2-D copy of bf16: dst[r, 0:cols] = src[r, 0:cols] */
int pt_copy2d_bf16(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int cols, void* stream);
/* nearest x2 along L (channels-last rows): y[b, 2l+{0,1}, :] = x[b, l, :]; bwd sums the pair */
int pt_upsample2_fwd(const void* x, void* y, int B, int L, int C, void* stream);
int pt_upsample2_bwd(const void* dy, void* dx, int B, int L, int C, void* stream);
/* [B, C, L] fp32 (reference layout) <-> [B, L, C] bf16 */
int pt_ncl_f32_to_nlc_bf16(const float* x, void* y, int B, int C, int L, void* stream);
int pt_nlc_bf16_to_ncl_f32(const void* x, float* y, int B, int C, int L, void* stream);
/* fp32 -> bf16 cast (Linear weights) and Conv1d weight repack [Co, Ci, k] fp32 -> [Co, k*Ci] bf16 */
int pt_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);
int pt_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream);
int pt_pack_conv_weight(const float* w, void* wp, int Co, int Ci, int k, void* stream);
/* gradient un-pack: g[Co, Ci, k] += gp[Co, k*Ci]  (fp32) */
int pt_unpack_conv_wgrad(const float* gp, float* g, int Co, int Ci, int k, int accumulate, void* stream);
/* column sums of a bf16 matrix: out[c] += sum_r x[r, c]  (bias gradients) */
int pt_colsum_bf16(const void* x, int64_t ld, float* out, int64_t rows, int cols, void* stream);
/* the same reduction in a 128-thread / 4 KB shape that fits beside a persistent GEMM CTA on the same SM (for launches on a second stream) */
int pt_colsum_bf16_lite(const void* x, int64_t ld, float* out, int64_t rows, int cols, void* stream);
/* per-(batch, channel) sum over L of dy[B, L, C] (time-shift gradient): out[b * out_stride + c] fp32 (overwritten) */
int pt_batch_colsum_bf16(const void* x, float* out, int64_t out_stride, int B, int L, int C, void* stream);

/* conv_in: x[B, L, Cin<=16] fp32 channels-last... k=3 pad 1 -> y[B, L, Co] bf16 (tts/ldm/unet_1d_condition.py:193,654) */
int pt_conv_in_fwd(const float* x_ncl, const float* w, const float* bias, void* y, int B, int Cin, int L, int Co, void* stream);
int pt_conv_in_bwd(const void* dy, const float* x_ncl, float* dw, float* dbias, int B, int Cin, int L, int Co, void* stream);
/* conv_out: h[B, L, C] bf16 -> y[B, Cout, L] fp32 (reference layout) (unet_1d_condition.py:410,734) */
int pt_conv_out_fwd(const void* h, const float* w, const float* bias, float* y_ncl, int B, int C, int L, int Cout, void* stream);
/* dy[B, Cout, L] fp32 -> dh[B, L, C] bf16 ; dw[Cout, C, 3], dbias[Cout] accumulated */
int pt_conv_out_bwd(const float* dy_ncl, const void* h, const float* w, void* dh, float* dw, float* dbias,
                    int B, int C, int L, int Cout, void* stream);

/* Timesteps (flip_sin_to_cos=True, shift 0): out[B, dim] fp32 = [cos | sin](t * exp(-ln(1e4) i / half)) */
int pt_time_sinusoid(const int64_t* t, float* out, int B, int dim, void* stream);
/* text front: y[b, l, :] = bf16(E[ids[b,l], :] + pe[l, :]) ; E fp32 [V, D], pe fp32 [L, D] */
int pt_text_embed_fwd(const int32_t* ids, const float* E, const float* pe, void* y, int B, int L, int D, int V, void* stream);
int pt_text_embed_bwd(const int32_t* ids, const void* dy, float* dE, int B, int L, int D, int V, void* stream);

/* train-step glue (train.py:96-98,107) */
int pt_add_noise(const float* x0, const float* noise, const int64_t* t, const float* sqrt_acp, const float* sqrt_1macp,
                 float* xt, int B, int64_t per_sample, void* stream);
/* One DDPM ancestral sampling step (diffusers 0.15 DDPMScheduler.step as the north-star sampling config uses it: epsilon
 * prediction, clip_sample, fixed_small variance; acp_t / acp_prev = alphas_cumprod at t and at the previous inference
 * timestep, acp_prev = 1 for the last step, which adds no noise).  x_prev = c_x0 * clamp(x0_hat, -1, 1) + c_xt * x_t + sigma * noise.
 * known != NULL: the first `keep` frames of every length-T row are overwritten from `known` (speech-prompt in-painting). */
int pt_ddpm_step(const float* eps, const float* xt, const float* noise, const float* known, float* out, int64_t n, int T, int keep,
                 float acp_t, float acp_prev, void* stream);
/* loss += mean((pred - target)^2) ; dpred = 2 (pred - target) / n * gscale ; loss must be zeroed by caller */
int pt_mse_fwd_bwd(const float* pred, const float* target, float* loss, float* dpred, int64_t n, float gscale, void* stream);

/* ------------------------------------------------------------------ RVQ ----------------------- */
/* codes[b, q, t] = argmin_j || r_q[b, :, t] - E[q, j, :] ||  (first index wins ties), r_{q+1} = r_q - E[q, code]
 * latents [B, D, T] fp32 (reference layout), codebooks [Q, K, D] fp32, codes [B, Q, T] int64.
 * Distances are evaluated exactly as the reference does: -(|r|^2 - 2 r.e + |e|^2) in fp32, the dot
 * product accumulated in ascending d order with fused multiply-add.
 * cb_sq: caller-provided scratch [Q, K] fp32 (|e|^2 per code, filled here) -- the library owns no device memory. */
int pt_rvq_encode_ws(const float* latents, const float* codebooks, float* cb_sq, int64_t* codes, int B, int D, int T, int Q, int K, void* stream);
int pt_rvq_cb_sq(const float* codebooks, float* out, int Q, int K, int D, void* stream);
/* The same codes (bit-identical, by construction) with the exhaustive fp32 search replaced by a tensor-core pre-selection: tcgen05
 * computes a bf16 approximation of every code's score, a rigorous error bound keeps the ~2 codes per frame and stage that can still be
 * the maximum, and only those are re-evaluated with the reference's exact fp32 arithmetic (csrc/rvq_tc.cu).  K % 128 == 0, K <= 1024.
 * scratch: pt_rvq_encode_tc_scratch_bytes(Q, K) bytes, 256-byte aligned, caller-owned; prep != 0 (re)derives the bf16 codebooks,
 * |e|^2 and max |e| from `codebooks` into it (needed once per codebook set). */
size_t pt_rvq_encode_tc_scratch_bytes(int Q, int K);
/* diagnostic: later pt_rvq_encode_tc launches dump the approximate stage-0 scores of their first 128 frames into buf[128][K]; NULL = off */
int pt_rvq_tc_debug_scores(float* buf);
int pt_rvq_encode_tc(const float* latents, const float* codebooks, void* scratch, int prep, int64_t* codes, int B, int D, int T, int Q, int K,
                     void* stream);
/* latents[b, :, t] = sum_q E[q, codes[b,q,t], :]  (q ascending, fp32) */
int pt_rvq_decode(const int64_t* codes, const float* codebooks, float* latents, int B, int D, int T, int Q, int K, void* stream);
/* same result (bit-equal), faster: scratch = caller-provided B*Q*T uint16 for the narrowed codes; a CTA keeps a 4-float slice of all
 * Q codebooks in shared memory instead of gathering 512-byte rows through L2.  Falls back to pt_rvq_decode if Q*K*16 B > 200 KB. */
int pt_rvq_decode_ws(const int64_t* codes, const float* codebooks, float* latents, void* scratch, int B, int D, int T, int Q, int K,
                     void* stream);
/* x0 = (codes / 1023 - 0.5) / 0.5  as fp32, and its inverse  codes = clamp(round((x + 1) * 511.5), 0, 1023) */
int pt_codes_affine(const int64_t* codes, float* x0, int64_t n, void* stream);
int pt_codes_affine_inv(const float* x, int64_t* codes, int64_t n, void* stream);

/* ------------------------------------------------------------------ optimiser --- */
/* sum of squares of a fp32 buffer accumulated into out[0] */
int pt_sumsq_f32(const float* x, int64_t n, float* out, void* stream);
/* AdamW step with gradient pre-scale  g *= min(1, max_norm / (sqrt(*gnorm_sq) + 1e-6)) * gscale  (train.py:41-47,116-120) */
int pt_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float wd, int step, const float* gnorm_sq, float max_norm, float gscale, void* stream);


/* Graph-safe variant: everything that changes between steps lives in device memory, so a captured step replays correctly.
 * `state` = PT_OPT_STATE_FLOATS fp32 words.  The host writes the hyper-parameter slots (an LR scheduler may rewrite PT_OPT_LR at any
 * time, outside the graph); pt_adamw_prepare increments the step counter (an int32 stored in slot PT_OPT_STEP), evaluates the bias
 * corrections in double precision and the clip factor min(1, max_norm / (sqrt(*gnorm_sq) * gscale + 1e-6)) * gscale, and leaves
 * them in the coefficient slots that pt_adamw_step_dev reads.  `g` is fp32, or bf16 when g_is_bf16 (gradients all-reduced in
 * bf16).  `w_bf16` (optional) receives bf16(p) in the same flat layout: the GEMM weight operands are views of it (no re-pack). */
enum {
  PT_OPT_LR = 0, PT_OPT_BETA1, PT_OPT_BETA2, PT_OPT_EPS, PT_OPT_WD, PT_OPT_MAX_NORM, PT_OPT_GSCALE, PT_OPT_STEP,
  PT_OPT_CLIP, PT_OPT_STEP_SIZE, PT_OPT_INV_SQRT_BC2, PT_OPT_DECAY, PT_OPT_GNORM, PT_OPT_GNORM_SQ, PT_OPT_STATE_FLOATS = 16
};
int pt_sumsq_bf16(const void* x, int64_t n, float* out, void* stream);
int pt_adamw_prepare(float* state, const float* gnorm_sq, void* stream);
/* deterministic global norm (bit-identical on every rank for identical gradients): `nparts` blocks write one partial sum of squares
 * each, pt_adamw_prepare_det adds them in a fixed order; the squared norm is left in slot PT_OPT_GNORM_SQ */
int pt_sumsq_partials(const void* x, int x_is_bf16, int64_t n, float* partials, int nparts, void* stream);
int pt_adamw_prepare_det(float* state, const float* partials, int nparts, void* stream);
int pt_adamw_step_dev(float* p, const void* g, int g_is_bf16, float* m, float* v, void* w_bf16, int64_t n, const float* state,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif
