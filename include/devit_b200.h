/*
 * devit_b200 -- C ABI of the B200-native DeViT collaborative-inference hot path.
 *
 * The reference (falcon-xu/DeViT) has no FFI layer: its "plugin boundary" for this path is
 * the set of nn.Module forwards in models/de_vit.py and models/ensemble_models.py.  Each entry
 * point below replaces the torch ops one of those forwards dispatches (file:line given per
 * function); the Python host (devit_b200/models.py) binds them with ctypes, see INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated; the caller owns all memory;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), no hidden syncs,
 *    safe to capture in a CUDA graph (the two entries that do synchronise say so:
 *    the one-off devit_pack_layer and the measurement hook devit_profile_collect);
 *  - return value 0 = ok, otherwise a DEVIT_ERR_* code; devit_last_error() gives the text;
 *  - there is no CPU fallback: on a device that is not sm_100 every launch returns
 *    DEVIT_ERR_DEVICE.
 *
 * Precision modes (`precision` fields)
 *  - DEVIT_BF16 : GEMM/attention operands are bf16 row-major arrays, fp32 accumulation in
 *                 TMEM, fp32 residual stream / LayerNorm statistics / softmax.
 *  - DEVIT_FP32 : operands are fp32 "split" arrays: two planes [hi | lo] with hi exactly
 *                 representable in tf32 and hi+lo == value; GEMMs run as 3xTF32 on tcgen05
 *                 (hi*hi + lo*hi + hi*lo), attention runs in plain fp32.  Matches the fp32
 *                 reference to ~1e-6 relative per op.
 */
#ifndef DEVIT_B200_H_
#define DEVIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEVIT_ABI_VERSION 4

enum {
  DEVIT_OK = 0,
  DEVIT_ERR_ARG = 1,     /* invalid argument (shape / alignment / null pointer) */
  DEVIT_ERR_DEVICE = 2,  /* current device is not compute capability 10.0 */
  DEVIT_ERR_CUDA = 3,    /* a CUDA runtime / driver call failed */
  DEVIT_ERR_WORKSPACE = 4 /* workspace too small */
};

enum { DEVIT_BF16 = 0, DEVIT_FP32 = 1 };

/* element kinds for outputs of the row-wise kernels and the GEMM epilogue */
enum {
  DEVIT_OUT_BF16 = 0,      /* bf16 [rows, ld] */
  DEVIT_OUT_F32 = 1,       /* fp32 [rows, ld] */
  DEVIT_OUT_F32_SPLIT = 2  /* fp32 hi plane at out, lo plane at out + plane_stride */
};

enum { DEVIT_ACT_NONE = 0, DEVIT_ACT_GELU_ERF = 1, DEVIT_ACT_RELU = 2 };

/* memory layout of uint8 image batches (devit_im2col_tokens_u8) */
enum { DEVIT_LAYOUT_NCHW = 0, DEVIT_LAYOUT_NHWC = 1 };

int devit_abi_version(void);
const char* devit_last_error(void);
/* 0 when the CURRENT cuda device is sm_100 (B200); DEVIT_ERR_DEVICE otherwise. */
int devit_device_check(void);
/* number of kernels this library has launched in the calling process (all threads) */
long long devit_launch_count(void);

/* Per-launch device timing for benchmarks.  devit_profile_enable(1) makes every launcher
 * bracket its kernel with two CUDA events on the launch stream (do not use while capturing a
 * CUDA graph); devit_profile_collect synchronises them and returns, per kernel family tag
 * (DEVIT_TAG_*, 16 slots), the summed milliseconds and the launch count, then clears the log. */
enum {
  DEVIT_TAG_GEMM_OTHER = 0, DEVIT_TAG_GEMM_PATCH = 1, DEVIT_TAG_GEMM_QKV = 2,
  DEVIT_TAG_GEMM_PROJ = 3, DEVIT_TAG_GEMM_FC1 = 4, DEVIT_TAG_GEMM_FC2 = 5,
  DEVIT_TAG_GEMM_FUSION = 6, DEVIT_TAG_GEMM_HEAD = 7, DEVIT_TAG_ATTENTION = 8,
  DEVIT_TAG_LAYERNORM = 9, DEVIT_TAG_GATHER_LN = 10, DEVIT_TAG_IM2COL = 11,
  DEVIT_TAG_PREFIX = 12, DEVIT_TAG_MLP_FUSED = 13, DEVIT_TAG_EVAL_TAIL = 14, DEVIT_NUM_TAGS = 16
};
int devit_profile_enable(int on);
int devit_profile_collect(double* ms_by_tag, long long* count_by_tag);

/* SM budget of the persistent kernels launched AFTER this call by any thread of the process
 * (0 = the whole chip, the default).  Every GEMM / attention / fused-MLP launch is a persistent
 * grid with one CTA per SM; when the host runs several independent kernel chains on separate
 * streams (the sub-models of an ensemble are independent until the fusion head,
 * models/ensemble_models.py:33) it sizes each chain's grids for a share of the chip so that two
 * chains run side by side and the HBM-bound phases of one overlap the tensor-bound phases of the
 * other.  Values are rounded down to an even count (CTA pairs); returns the previous budget. */
int devit_set_sm_budget(int sms);

/* The launchers cache the TMA descriptors (CUtensorMap) they encode, keyed by base pointer +
 * dims + strides + box + swizzle, behind a mutex (bounded; DEVIT_TMAP_CACHE=0 switches it off).
 * Returns the number of cached descriptors; *hits / *misses (either may be NULL) receive the
 * process-wide lookup counters. */
int devit_tmap_cache_stats(long long* hits, long long* misses);

/* ---------------------------------------------------------------------------------------
 * devit_gemm: out = epilogue( sum_s A_s[M, K_s] * B_s[N, K_s]^T )  on tcgen05/TMEM via TMA.
 *
 * Replaces nn.Linear (addmm) call sites: qkv  models/de_vit.py:67, proj :81, fc1 :36,
 * fc2 :45, head/head_dist :317, patch-embed conv-as-GEMM (timm PatchEmbed, called :258),
 * EnsMLP cls_mlp/dist_mlp/cls_classifier/dist_classifier models/ensemble_models.py:79-85.
 *
 * A and B are both "K-major" (row-major with K contiguous): A is [a_rows, a_cols] with row
 * stride lda, B (an nn.Linear weight) is [b_rows, b_cols] with row stride ldb.  Up to 8
 * K-segments are accumulated: segment s multiplies A rows [a_row_off + m] cols
 * [a_k_off, a_k_off + k_len) with B cols [b_k_off, b_k_off + k_len)  (this is how the fusion
 * head consumes the all-gathered [n_sub, B, D] feature slabs without a stack/transpose).
 * Reads outside [a_rows, a_cols] / [b_rows, b_cols] are zero-filled by TMA, so M, N and the
 * last K block may be ragged.
 *
 * Epilogue, per output element (m, n), all in fp32:
 *   v = acc + bias[n]; v = act(v); v += rowbias[(rowmap_off + m % period) * ld_rowbias + n];
 *   v += resid[row_out * ldr + n]; v *= alpha; out[row_out, n] = v
 *   row_out = period ? (m / period) * rowmap_stride + rowmap_off + m % period : m
 * (each term is skipped when its pointer is NULL; resid may alias out when out is fp32).
 * ------------------------------------------------------------------------------------- */
typedef struct devit_gemm_seg {
  int32_t a_row_off;
  int32_t a_k_off;
  int32_t b_k_off;
  int32_t k_len;
} devit_gemm_seg;

typedef struct devit_gemm_args {
  int32_t precision; /* DEVIT_BF16 | DEVIT_FP32 (operand format) */
  int32_t m, n;
  const void* a;
  int32_t a_rows, a_cols;
  int64_t lda;
  int64_t a_plane_stride; /* elements between hi and lo plane (DEVIT_FP32 only) */
  const void* b;
  int32_t b_rows, b_cols;
  int64_t ldb;
  int64_t b_plane_stride;
  int32_t num_segs;
  devit_gemm_seg segs[8];
  void* out;
  int64_t ldo;
  int32_t out_kind; /* DEVIT_OUT_* */
  int64_t out_plane_stride;
  const float* bias;
  const float* resid;
  int64_t ldr;
  int32_t resid_period; /* 0: resid is [m, n].  > 0: resid is a TABLE of resid_period + 31 rows and
                           output row r adds table row r % resid_period; rows [period, period+31)
                           repeat rows [0, 31).  This is how pos_embed (+ cls/dist) is added by the
                           patch GEMM without materialising it per image.  fp32 outputs only. */
  const float* rowbias;
  int64_t ld_rowbias;
  int32_t act;
  float alpha;
  int32_t rowmap_period, rowmap_stride, rowmap_off;
  int32_t block_n; /* 0 = choose automatically from {128, 192, 256} */
  int32_t profile_tag; /* DEVIT_TAG_GEMM_* bucket used by devit_profile_collect */
  int32_t cluster_m;   /* 1 = one CTA per 128-row tile; 2 = CTA pair (cta_group::2) per
                          256-row tile, each CTA staging half of the weight tile; 0 = auto */
  /* ---- LayerNorm folding (DEVIT_BF16 operands only; all NULL/0 = off).
   * A LayerNorm followed by a Linear (norm1 -> qkv, models/de_vit.py:113 + :67; norm2 -> fc1,
   * :115 + :36) is computed WITHOUT materialising the normalised tensor:
   *     LN(x) W^T + b = rstd_m * (x (gamma .* W)^T - mean_m * c1) + c2,
   *     c1[n] = sum_k gamma_k W[n,k],  c2[n] = b[n] + sum_k beta_k W[n,k].
   * Consumer GEMM (bf16 output): `a` is the raw residual stream rounded to bf16, `b` holds the
   * gamma-folded weights, `bias` holds c2, `ln_colsum` holds c1 and `ln_stats` the per-row
   * partial sums [ln_parts][m][2] = (sum x, sum x^2) over disjoint column ranges of the fp32
   * stream (1 <= ln_parts <= 12); the epilogue derives mean / rstd (biased variance over ln_dim, + ln_eps) per row.
   * Producer GEMM (fp32 output with `resid`, n % 128 == 0): additionally writes the bf16 copy
   * of its output to `out_bf16` and the partial row sums of its output to `stats_out`
   * [2 * n / 128][m][2] (one part per 64 output columns). */
  const float* ln_stats;
  int32_t ln_parts;
  int32_t ln_dim;
  float ln_eps;
  const float* ln_colsum;
  void* out_bf16;
  int64_t ld_out_bf16;
  float* stats_out;
} devit_gemm_args;

int devit_gemm(const devit_gemm_args* args, void* stream);

/* Debug aid: device buffer of 20 x 512 int64 that receives clock64 stamps of the pipeline
 * roles of GEMM CTA 0 (producer / MMA issuer / epilogue warp 0); NULL switches it off. */
int devit_debug_set_trace(long long* device_buf);

/* ---------------------------------------------------------------------------------------
 * devit_layernorm: y[r, :] = (x[r, :] - mean) * rstd * gamma + beta, biased variance,
 * one warp per row, fp32 statistics.   Replaces nn.LayerNorm (norm1/norm2)
 * models/de_vit.py:113,115.  x fp32 [rows, dim] (the residual stream), dim in {256,384,768}.
 * ------------------------------------------------------------------------------------- */
int devit_layernorm(const float* x, const float* gamma, const float* beta, void* y,
                    int64_t rows, int32_t dim, float eps, int32_t out_kind,
                    int64_t out_plane_stride, void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_mlp_fused:  x += gelu( LN(x) W1^T + b1 ) W2^T + b2   as ONE kernel (DEVIT_BF16, dim 384
 * or 256).
 * Replaces norm2 + Mlp.forward + the residual add, models/de_vit.py:35-47 and :115: fc1, the
 * erf-GELU, the (compacted) neuron gate and fc2 run per CTA pair on 256 token rows with the
 * hidden activation kept in tensor memory, so the [rows, hidden] tensor never exists in HBM.
 * LayerNorm is folded exactly as in devit_gemm_args.ln_stats: `xb` is the bf16 copy of the fp32
 * residual stream `x`, `w1` holds gamma-folded weights [hidden_ld, dim], `c1` / `c2` the column
 * sums / folded bias [hidden_ld] (zero beyond the kept neurons), `ln_stats` the partial row sums
 * [ln_parts][m][2].  `w2` is [dim, hidden_ld] (zero columns beyond the kept neurons).
 * Outputs: x (fp32, in place), optionally the bf16 copy of the new x (`xb_out`, may alias `xb`)
 * and its partial row sums `stats_out` [4][m][2] (one part per dim/4 columns) for the next layer.
 * Optionally the attention-output projection is fused in front (fields `o` .. `proj_k` below).
 * ------------------------------------------------------------------------------------- */
typedef struct devit_mlp_args {
  int32_t m;
  int32_t dim;       /* 384 or 256 */
  int32_t hidden_ld; /* kept neurons rounded up to a multiple of 16 */
  const void* xb;
  const void* w1;
  const float* c1;
  const float* c2;
  const float* ln_stats;
  int32_t ln_parts;
  float ln_eps;
  const void* w2;
  const float* b2;
  float* x;
  void* xb_out;
  float* stats_out;
  /* ---- optional: the attention-output projection fused in front (all NULL / 0 = off).
   * With `o` set the kernel computes
   *     x1 = x + o Wp^T + bp ;   x = x1 + gelu( LN(x1) W1^T + b1 ) W2^T + b2
   * i.e. Attention.proj + the first residual add (models/de_vit.py:81-82, :114) and the whole MLP
   * branch (:35-47, :115) in one pass over the residual stream: x1, its bf16 copy and its
   * LayerNorm statistics never leave the SM pair (`xb`, `ln_stats`, `ln_parts` are ignored; the
   * exact row statistics of x1 are formed on chip).
   * o: attention output [m, proj_k] bf16 (proj_k = 64 * kept heads, <= dim), w_proj: [dim, proj_k]
   * bf16 (kept heads' columns), b_proj: fp32 [dim]. */
  const void* o;
  const void* w_proj;
  const float* b_proj;
  int32_t proj_k;
  /* ---- optional (with `o` only): the residual stream as TWO bf16 planes, x = hi + lo
   * (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits, 2^-17 relative), instead of fp32 x plus a
   * bf16 copy.  An SM stores at most ~32 bytes per clock, so 4 bytes per element instead of 6
   * shorten the kernel's write-bound final epilogue by a third; hi is at the same time the A
   * operand of the next layer's LayerNorm-folded QKV GEMM.
   * x_lo_in  != NULL: the input residual is `xb` (hi plane) + x_lo_in; fp32 `x` is not read.
   * x_lo_out != NULL: the output is written as `xb_out` (hi plane) + x_lo_out; fp32 `x` is not
   *                   written (the last layer leaves it NULL so the final norm reads fp32 x).
   * Planes are bf16 [m, dim], 16-byte aligned; in and out planes may alias (tile-local). */
  const void* x_lo_in;
  void* x_lo_out;
} devit_mlp_args;

int devit_mlp_fused(const devit_mlp_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_rowstats: xb[r, :] = bf16(x[r, :]) and stats[r] = (sum_k x[r,k], sum_k x[r,k]^2): the
 * one-part input of a LayerNorm-folded GEMM (see devit_gemm_args.ln_stats) for a residual
 * stream that was not produced by a devit_gemm (the token embedding, models/de_vit.py:258-264).
 * x fp32 [rows, dim], dim in {256,384,768}; one warp per row.
 * ------------------------------------------------------------------------------------- */
int devit_rowstats(const float* x, void* xb, float* stats, int64_t rows, int32_t dim,
                   void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_attention: out[b, t, h*64:(h+1)*64] = softmax(q k^T * scale) v   per (b, h).
 * Replaces models/de_vit.py:70-74 (two bmm + scale + softmax + transpose).
 *
 * qkv is the natural output of the QKV Linear: [batch*tokens, 3*heads*64] with column
 * (which*heads + h)*64 + d  (the reference's reshape(B,N,3,H,hd), models/de_vit.py:67);
 * out is [batch*tokens, heads*64] -- exactly the A operand of the proj GEMM, so the
 * reference's transpose(1,2).reshape copies vanish.  `heads` is the number of KEPT heads
 * (gated heads are compacted away by the host).  head_dim is fixed at 64, tokens <= 256.
 * DEVIT_BF16: bf16 in/out, QK^T and PV on tcgen05, fp32 softmax.
 * DEVIT_FP32: split-fp32 in (hi+lo summed on load) and split-fp32 out, fp32 CUDA-core math.
 * ------------------------------------------------------------------------------------- */
int devit_attention(int32_t precision, const void* qkv, int64_t qkv_plane_stride, void* out,
                    int64_t out_plane_stride, int32_t batch, int32_t tokens, int32_t heads,
                    float scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_im2col_patch16: images fp32 NCHW [batch, chans, hw, hw] -> patch matrix
 * A[batch*(hw/16)^2, chans*256] with K order (c, py, px) and patch index gy*(hw/16)+gx, i.e.
 * the operand that turns timm PatchEmbed's Conv2d(k=16, s=16) + flatten(2).transpose(1,2)
 * (called at models/de_vit.py:258) into a GEMM against proj.weight.view(D, chans*256).
 * ------------------------------------------------------------------------------------- */
int devit_im2col_patch16(const float* images, void* a, int32_t batch, int32_t chans,
                         int32_t hw, int32_t out_kind, int64_t out_plane_stride, void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_im2col_tokens / devit_token_init: the same patch embedding laid out so that it is a plain
 * residual GEMM over TOKEN rows.  im2col_tokens writes A[batch * tokens, chans*256] with
 * tokens = num_prefix + (hw/16)^2: the first num_prefix rows of every image are zero, row
 * num_prefix + p holds patch p.  token_init writes x[b, j, :] = pos[j, :] + (j < num_prefix ?
 * prefix[j, :] - bias : 0), so that  x += A W^T + bias  yields cls/dist + pos on the prefix rows
 * and patch embedding + pos on the others (timm PatchEmbed + models/de_vit.py:258-264) through the
 * coalesced TMA epilogue, and A can be shared by every sub-model that sees the same images.
 * ------------------------------------------------------------------------------------- */
int devit_im2col_tokens(const float* images, void* a, int32_t batch, int32_t chans, int32_t hw,
                        int32_t num_prefix, int32_t out_kind, int64_t out_plane_stride,
                        void* stream);
int devit_token_init(float* x, const float* prefix, const float* pos, const float* bias,
                     int32_t batch, int32_t tokens, int32_t dim, int32_t num_prefix, void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_im2col_tokens_u8: the same token-row patch matrix as devit_im2col_tokens, computed from
 * the DECODED uint8 images, with the input normalisation applied on the way:
 *     v = ((float)u8 / 255 - mean[c]) / std[c]      (fp32, IEEE division, no contraction)
 * i.e. torchvision ToTensor + Normalize of the reference's eval transform
 * (data/get_dataset.py:107-108) and images.to(device) (engine.py:224) folded into the patch
 * extraction: a batch crosses PCIe as bytes (4x less than fp32) and the fp32 image tensor is never
 * materialised.  The fp32 values are bit-identical to the CPU transform's.
 * images: DEVIT_LAYOUT_NCHW [batch, chans, hw, hw] or DEVIT_LAYOUT_NHWC [batch, hw, hw, 3];
 * mean / stdv: HOST pointers to `chans` floats (1 <= chans <= 4).
 * ------------------------------------------------------------------------------------- */
int devit_im2col_tokens_u8(const uint8_t* images, int32_t layout, const float* mean,
                           const float* stdv, void* a, int32_t batch, int32_t chans, int32_t hw,
                           int32_t num_prefix, int32_t out_kind, int64_t out_plane_stride,
                           void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_eval_tail: the per-batch tail of engine.evaluate / evaluate_ens_disjoint
 * (engine.py:229-238): CrossEntropyLoss(logits, target) (mean over the batch) and timm
 * accuracy(logits, target, topk=(1, k)), accumulated ON THE DEVICE so the evaluation loop needs
 * one host synchronisation per epoch instead of three .item() calls per batch.
 *   logits fp32 [batch, classes] (row stride ld), target int64 [batch];
 *   acc (optional) double[5], updated in place:  acc[0] += mean loss of this batch, acc[1] += 1,
 *       acc[2] += #correct@1, acc[3] += #correct@k, acc[4] += batch   -- exactly the totals /
 *       counts the reference's MetricLogger meters receive (utils/dist_utils.py:30-33);
 *   batch_out (optional) float[3] = {mean loss, #correct@1, #correct@k} of this batch;
 *   workspace: devit_eval_tail_workspace_bytes(batch) bytes.
 * A sample counts as correct@k when fewer than k classes sort before its target in descending
 * logit order (ties: the smaller class index first).  A target outside [0, classes) gives a NaN
 * loss and is never correct.  Deterministic (fixed reduction order, no float atomics).
 * ------------------------------------------------------------------------------------- */
size_t devit_eval_tail_workspace_bytes(int32_t batch);
int devit_eval_tail(const float* logits, int64_t ld, const int64_t* target, int32_t batch,
                    int32_t classes, int32_t topk, void* workspace, size_t workspace_bytes,
                    double* acc, float* batch_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_token_prefix: x[b, j, :] = prefix[j, :] + pos[j, :] for j < num_prefix
 * (cls / dist tokens; models/de_vit.py:259-264).  x fp32 [batch, tokens, dim].
 * ------------------------------------------------------------------------------------- */
int devit_token_prefix(float* x, const float* prefix, const float* pos, int32_t batch,
                       int32_t tokens, int32_t dim, int32_t num_prefix, void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_gather_ln: feats[j, b, :] = LayerNorm(x[b, j, :]) for j < num_prefix -- the final
 * norm restricted to the rows the model actually returns (models/de_vit.py:286-288).
 * feats_f32 (optional) fp32 [num_prefix, kind_rows, dim]; feats_op (optional) the same values in
 * the GEMM operand format selected by out_kind (bf16 or split fp32) for the fusion head.
 * kind_rows (0 = batch) is the row count of one token kind in the OUTPUT slabs: a caller that
 * runs a batch in several chunks passes the slab pointers advanced to the chunk's first image and
 * kind_rows = the whole batch.
 * ------------------------------------------------------------------------------------- */
int devit_gather_ln(const float* x, const float* gamma, const float* beta, float* feats_f32,
                    void* feats_op, int32_t out_kind, int64_t out_plane_stride, int32_t batch,
                    int32_t tokens, int32_t dim, int32_t num_prefix, float eps,
                    int32_t kind_rows, void* stream);

/* ---------------------------------------------------------------------------------------
 * devit_vit_forward: the whole VisionTransformer.forward_features of one (compacted)
 * sub-model, models/de_vit.py:242-292: patch embed -> tokens -> depth x Block -> final norm on
 * the cls/dist rows.  One call enqueues every kernel on `stream`.
 *
 * Weights are the host-packed, gate-compacted operands (devit_b200/packing.py): per layer the
 * kept heads' qkv rows / proj columns and the kept neurons' fc1 rows / fc2 columns.
 * ------------------------------------------------------------------------------------- */
typedef struct devit_layer_desc {
  int32_t heads;     /* kept heads h_l (>= 1)                                  */
  int32_t hidden;    /* kept neurons f_l                                        */
  int32_t hidden_ld; /* f_l rounded up to a multiple of 16 (row stride of fc2 / hidden)*/
  const float* ln1_g;
  const float* ln1_b;
  const void* w_qkv; /* [3*h_l*64, dim]   */
  const float* b_qkv;
  const void* w_proj; /* [dim, h_l*64]     */
  const float* b_proj;
  const float* ln2_g;
  const float* ln2_b;
  const void* w_fc1; /* [hidden_ld, dim] (rows >= hidden are zero) */
  const float* b_fc1; /* [hidden_ld]       */
  const void* w_fc2; /* [dim, hidden_ld] (cols >= hidden are zero) */
  const float* b_fc2;
  /* LayerNorm folding (DEVIT_BF16 only, both NULL = off): when set, w_qkv / w_fc1 hold the
   * gamma-folded weights, b_qkv / b_fc1 hold c2 and these hold c1 (devit_gemm_args.ln_colsum);
   * ln1/ln2 are then not read and no normalised tensor is ever written. */
  const float* cs_qkv; /* [3*h_l*64] */
  const float* cs_fc1; /* [hidden_ld] */
} devit_layer_desc;

/* ---------------------------------------------------------------------------------------
 * devit_pack_layer: one Block's fp32 master parameters + head / neuron gates -> the
 * gate-compacted (and, optionally, LayerNorm-folded) operands of a devit_layer_desc.
 *
 * This is the device-side form of the weight preparation the drop-in modules need whenever a
 * gate (core/imp_rank.py:65-71, :147-153 assign `m.gate`) or a parameter changes: kept heads
 * (gate != 0, ascending) select the q/k/v rows of qkv.weight / bias and the columns of
 * proj.weight (scaled by the gate value); kept neurons select rows of fc1.weight / bias and
 * columns of fc2.weight (scaled), zero-padded to a multiple of 16.  With fold_ln (DEVIT_BF16
 * only) qkv / fc1 weights are multiplied by the preceding LayerNorm's gamma, `cs_*` receive the
 * row sums of the bf16-rounded folded weights and `b_qkv` / `b_fc1` receive b + W beta
 * (devit_gemm_args.ln_stats).  If every head is gated off, head 0 is kept with zeroed proj
 * columns.  All parameter pointers are DEVICE pointers (b_qkv may be NULL); the gates are HOST
 * arrays (NULL = all ones).  `packed` (device, 256-byte aligned, devit_pack_layer_bytes bytes,
 * owned by the caller) receives every array `out` points at, except ln1/ln2/b_proj/b_fc2, which
 * alias the inputs.  kept_heads / kept_neurons (HOST, optional, sized num_heads / hidden) receive
 * the kept indices.  Synchronises `stream` before returning (one-off preparation work).
 * ------------------------------------------------------------------------------------- */
typedef struct devit_block_weights {
  int32_t dim, num_heads, hidden;
  const float* ln1_g;
  const float* ln1_b;
  const float* w_qkv; /* [3*dim, dim] */
  const float* b_qkv; /* [3*dim] or NULL */
  const float* w_proj; /* [dim, dim] */
  const float* b_proj;
  const float* ln2_g;
  const float* ln2_b;
  const float* w_fc1; /* [hidden, dim] */
  const float* b_fc1;
  const float* w_fc2; /* [dim, hidden] */
  const float* b_fc2;
  const float* head_gate;   /* HOST [num_heads] or NULL */
  const float* neuron_gate; /* HOST [hidden] or NULL */
} devit_block_weights;

size_t devit_pack_layer_bytes(const devit_block_weights* w, int32_t precision);
int devit_pack_layer(const devit_block_weights* w, int32_t precision, int32_t fold_ln,
                     void* packed, size_t packed_bytes, devit_layer_desc* out,
                     int32_t* kept_heads, int32_t* num_kept_heads, int32_t* kept_neurons,
                     int32_t* num_kept_neurons, void* stream);

typedef struct devit_vit_desc {
  int32_t precision; /* DEVIT_BF16 | DEVIT_FP32 */
  int32_t dim;       /* 384 (dedeit) or 768 (teacher); head_dim = 64 */
  int32_t depth;
  int32_t img, chans;  /* 224, 3 ; patch size fixed at 16 */
  int32_t num_prefix;  /* 2 = cls + dist (distilled), 1 = cls only */
  float ln_eps;
  const void* w_patch; /* [dim, chans*256] */
  const float* b_patch;
  const float* prefix; /* [num_prefix, dim] = cls_token (, dist_token) */
  const float* pos;    /* [num_prefix + (img/16)^2, dim] */
  const float* norm_g;
  const float* norm_b;
  const devit_layer_desc* layers; /* HOST pointer to depth entries */
  int64_t w_plane_stride_unused;  /* reserved, must be 0 */
  /* Optional [tokens + 31, dim] fp32 table: row j < tokens = pos[j] + (j < num_prefix ?
   * prefix[j] - b_patch : 0), rows [tokens, tokens + 31) repeat rows [0, 31).  When set, the
   * patch GEMM adds it as a periodic residual (devit_gemm_args.resid_period = tokens) and
   * devit_token_init is not launched. */
  const float* tok_table;
} devit_vit_desc;

size_t devit_vit_workspace_bytes(const devit_vit_desc* desc, int32_t batch);

/* feats_f32: fp32 [num_prefix, batch, dim]; feats_op: GEMM-operand copy (bf16, or split fp32
 * with the given plane stride) or NULL.  x_out (optional): fp32 [batch, tokens, dim] copy of
 * the residual stream after the last block that ran; num_layers_run < 0 runs all `depth`
 * blocks, otherwise only the first num_layers_run (parity checks of intermediate state). */
int devit_vit_forward(const devit_vit_desc* desc, const float* images, int32_t batch,
                      void* workspace, size_t workspace_bytes, float* feats_f32, void* feats_op,
                      int64_t feats_op_plane_stride, float* x_out, int32_t num_layers_run,
                      void* stream);

/* Same, starting from the token-row patch matrix `patches` (devit_im2col_tokens output in the
 * descriptor's operand format; for DEVIT_FP32 the lo plane follows at patches_plane_stride
 * elements) instead of the images: the N sub-models of an ensemble all embed the SAME images
 * (models/ensemble_models.py:33), so the host extracts the patches once per batch. */
int devit_vit_forward_patches(const devit_vit_desc* desc, const void* patches,
                              int64_t patches_plane_stride, int32_t batch, void* workspace,
                              size_t workspace_bytes, float* feats_f32, void* feats_op,
                              int64_t feats_op_plane_stride, float* x_out, int32_t num_layers_run,
                              void* stream);

/* Same as devit_vit_forward / devit_vit_forward_patches (pass exactly one of `images` /
 * `patches`) with per-layer exports for the training-side consumers of this forward
 * (VisionTransformer.forward(..., output_qkv=True), models/de_vit.py:268-284; the teacher's
 * q/k/v feed feature_relation_loss, engine.py:70-95, utils/losses.py:307-327):
 * exports->qkv[l] (or NULL) receives layer l's QKV Linear output [batch * tokens, 3 * heads_l * 64]
 * in the descriptor's operand format, column (which * heads_l + h) * 64 + d (the reference's
 * reshape(B, N, 3, H, hd), models/de_vit.py:67) -- bf16, or fp32 hi plane followed by the lo
 * plane at + batch * tokens * 3 * heads_l * 64 elements.  The GEMM writes there directly and the
 * attention kernel reads it back, so an export costs no copy.  Only KEPT heads exist in a
 * compacted sub-model; the host uses this path when no head is gated off. */
#define DEVIT_MAX_DEPTH 32
typedef struct devit_vit_exports {
  void* qkv[DEVIT_MAX_DEPTH];
  /* rows of one token kind in feats_f32 / feats_op (0 = batch): set to the whole batch when this
   * call processes a chunk of it and the feats pointers address the chunk's first image */
  int32_t feats_kind_rows;
} devit_vit_exports;

int devit_vit_forward_ex(const devit_vit_desc* desc, const float* images, const void* patches,
                         int64_t patches_plane_stride, int32_t batch, void* workspace,
                         size_t workspace_bytes, float* feats_f32, void* feats_op,
                         int64_t feats_op_plane_stride, float* x_out, int32_t num_layers_run,
                         const devit_vit_exports* exports, void* stream);

/* ---------------------------------------------------------------------------------------
 * CCT (Compact Convolutional Transformer) sub-models: models/cct.py:138-157 +
 * models/utils/tokenizer.py:23-44 + models/utils/transformers.py:104-113, :441-477.
 *   tokenizer : n_conv x [conv k3 s1 p1 (no bias) -> ReLU -> maxpool 3/2/1], each conv as
 *               im2col (K order (ky, kx, c), zero-padded to a multiple of 8) + devit_gemm with a
 *               ReLU epilogue, then a channels-last max-pool; the last pool adds positional_emb
 *               and writes the fp32 residual stream [batch * tokens, dim];
 *   blocks    : the same pre-norm blocks as the ViT path (QKV without bias, LN eps 1e-5);
 *   head      : LayerNorm over every token, sequence pooling
 *               pooled[b] = softmax_t(x[b,t,:] . pool_w + pool_b)^T x[b]   (:473).
 * `pooled` is fp32 [batch, dim] (the backbone output / the input of the fc or fusion GEMM).
 * ------------------------------------------------------------------------------------- */
typedef struct devit_cct_desc {
  int32_t precision;   /* DEVIT_BF16 | DEVIT_FP32 */
  int32_t dim;         /* 256 (cct_7) */
  int32_t depth;
  int32_t img, chans;  /* 32, 3 */
  int32_t n_conv;      /* 1 or 2 */
  int32_t conv_chans[3]; /* output channels of conv layer i (last == dim) */
  const void* w_conv[3]; /* [conv_chans[i], kpad_i] operand format, K order (ky, kx, c_in) */
  int32_t conv_kpad[3];  /* 9 * c_in rounded up to a multiple of 8 */
  const float* pos;    /* [tokens, dim] or NULL */
  float ln_eps;
  const float* norm_g;
  const float* norm_b;
  const float* pool_w; /* [dim] */
  float pool_b;
  const devit_layer_desc* layers; /* HOST pointer; b_qkv may be NULL (no QKV bias) */
} devit_cct_desc;

size_t devit_cct_workspace_bytes(const devit_cct_desc* desc, int32_t batch);
int devit_cct_forward(const devit_cct_desc* desc, const float* images, int32_t batch,
                      void* workspace, size_t workspace_bytes, float* pooled, float* x_out,
                      int32_t num_layers_run, void* stream);

/* Building blocks of the above (exposed for the parity tests). */
int devit_im2col3x3(const float* in, void* a, int32_t batch, int32_t chans, int32_t hw,
                    int64_t stride_b, int64_t stride_c, int64_t stride_y, int64_t stride_x,
                    int32_t kpad, int32_t out_kind, int64_t out_plane_stride, void* stream);
int devit_maxpool3x3s2_cl(const void* in, int32_t in_kind, float* out, const float* pos,
                          int32_t batch, int32_t hw, int32_t chans, void* stream);
int devit_seqpool(const float* xn, const float* w, float b, float* pooled, int32_t batch,
                  int32_t tokens, int32_t dim, void* stream);

/* In DEVIT_FP32 every weight matrix pointer in the descriptors addresses the hi plane and the
 * lo plane follows at + rows*cols elements (plane stride = rows * ld). */

#ifdef __cplusplus
}
#endif
#endif /* DEVIT_B200_H_ */
