// Whole-sub-model forward (VisionTransformer.forward_features, models/de_vit.py:242-292) as one
// C call that enqueues every kernel on the caller's stream: im2col -> patch GEMM (+bias +pos,
// row-remapped into the token matrix) -> cls/dist rows -> depth x [LN, QKV GEMM, attention,
// proj GEMM (+residual), LN, fc1 GEMM (+GELU), fc2 GEMM (+residual)] -> final LN on the
// cls/dist rows only.  Gated heads / neurons never appear: the host packs compacted weights.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace devit {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct VitLayout {
  int tokens, grid, max_heads, max_hidden_ld, planes, esz;
  long long M;
  size_t off_x, off_y, off_qkv, off_o, off_hid, off_stats, total;
  int stat_parts;  // partial row sums written by a LayerNorm-folding producer GEMM
};

static int plan_layout(const devit_vit_desc* d, int batch, VitLayout* L) {
  DEVIT_REQUIRE(d && d->layers, "devit_vit: null descriptor");
  DEVIT_REQUIRE(d->precision == DEVIT_BF16 || d->precision == DEVIT_FP32,
                "devit_vit: bad precision %d", d->precision);
  DEVIT_REQUIRE(d->dim == 256 || d->dim == 384 || d->dim == 768,
                "devit_vit: dim %d not in {256,384,768}", d->dim);
  DEVIT_REQUIRE(d->depth > 0 && d->img > 0 && d->img % 16 == 0 && d->chans > 0 && batch > 0,
                "devit_vit: bad geometry");
  DEVIT_REQUIRE(d->num_prefix == 1 || d->num_prefix == 2, "devit_vit: num_prefix must be 1 or 2");
  L->grid = d->img / 16;
  L->tokens = L->grid * L->grid + d->num_prefix;
  DEVIT_REQUIRE(L->tokens <= 256, "devit_vit: %d tokens exceed the attention kernel's 256",
                L->tokens);
  L->M = static_cast<long long>(batch) * L->tokens;
  L->planes = d->precision == DEVIT_BF16 ? 1 : 2;
  L->esz = d->precision == DEVIT_BF16 ? 2 : 4;
  L->max_heads = 0;
  L->max_hidden_ld = 0;
  for (int l = 0; l < d->depth; ++l) {
    const devit_layer_desc& y = d->layers[l];
    DEVIT_REQUIRE(y.heads >= 1 && y.heads * 64 <= d->dim, "devit_vit: layer %d heads %d", l,
                  y.heads);
    DEVIT_REQUIRE(y.hidden >= 1 && y.hidden_ld >= y.hidden && y.hidden_ld % 16 == 0,
                  "devit_vit: layer %d hidden %d / ld %d (ld must be a multiple of 16)", l,
                  y.hidden, y.hidden_ld);
    if (y.heads > L->max_heads) L->max_heads = y.heads;
    if (y.hidden_ld > L->max_hidden_ld) L->max_hidden_ld = y.hidden_ld;
  }
  const size_t pe = static_cast<size_t>(L->planes) * L->esz;
  size_t off = 0;
  L->off_x = off;
  off = align_up(off + static_cast<size_t>(L->M) * d->dim * 4, 256);
  L->off_y = off;
  off = align_up(off + static_cast<size_t>(L->M) * d->dim * pe, 256);
  L->off_qkv = off;
  const size_t qkv_b = static_cast<size_t>(L->M) * 3 * L->max_heads * 64 * pe;
  const size_t patch_b = static_cast<size_t>(L->M) * d->chans * 256 * pe;  // token-row patches
  off = align_up(off + (qkv_b > patch_b ? qkv_b : patch_b), 256);
  L->off_o = off;
  off = align_up(off + static_cast<size_t>(L->M) * L->max_heads * 64 * pe, 256);
  L->off_hid = off;
  // (also holds the lo plane of the split residual stream: at least dim columns)
  off = align_up(off + static_cast<size_t>(L->M) *
                           (L->max_hidden_ld > d->dim ? L->max_hidden_ld : d->dim) * pe, 256);
  L->off_stats = off;
  L->stat_parts = 2 * ((d->dim + 127) / 128);
  off = align_up(off + static_cast<size_t>(L->M) * L->stat_parts * 2 * sizeof(float), 256);
  L->total = off;
  return DEVIT_OK;
}

// DEVIT_SYNC_DEBUG=1: synchronise and report after every op of the forward (debug aid)
static int sync_debug(const char* what, int layer, void* stream) {
  static int on = -1;
  if (on < 0) on = getenv("DEVIT_SYNC_DEBUG") ? 1 : 0;
  if (!on) return DEVIT_OK;
  cudaError_t e = cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
  fprintf(stderr, "[devit] layer %d %s: %s\n", layer, what, cudaGetErrorString(e));
  fflush(stderr);
  if (e != cudaSuccess) return set_error(DEVIT_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
  return DEVIT_OK;
}

static void base_gemm(devit_gemm_args* g, int precision) {
  std::memset(g, 0, sizeof(*g));
  g->precision = precision;
  g->num_segs = 1;
  g->alpha = 1.0f;
}

}  // namespace devit

using namespace devit;

extern "C" size_t devit_vit_workspace_bytes(const devit_vit_desc* desc, int32_t batch) {
  VitLayout L{};
  if (plan_layout(desc, batch, &L)) return 0;
  return L.total;
}

// The transformer blocks (models/de_vit.py:103-121; models/utils/transformers.py:104-113 has the
// same pre-norm structure) over a residual stream x [M, D] that already holds the embedded
// tokens.  In `fold` mode y / stats must hold the bf16 copy of x and `parts` partial row sums.
struct BlockBuffers {
  float* x;
  void* y;
  void* qkv;
  void* o;
  void* hid;
  float* stats;
  void* const* qkv_export = nullptr;  // per layer: caller buffer that receives q/k/v (or NULL)
};

static int run_blocks(const devit_layer_desc* layers, int nl, int prec, int D, float ln_eps,
                      int batch, int tokens, long long M, const BlockBuffers& bufs, bool fold,
                      int parts, int stat_parts, void* stream) {
  int rc = DEVIT_OK;
  const int opk = prec == DEVIT_BF16 ? DEVIT_OUT_BF16 : DEVIT_OUT_F32_SPLIT;
  float* x = bufs.x;
  void* y = bufs.y;
  void* qkv = bufs.qkv;
  void* o = bufs.o;
  void* hid = bufs.hid;
  float* stats = bufs.stats;
  devit_gemm_args g;
  static int fused_mlp = -1;  // DEVIT_FUSED_MLP=0: separate fc1 / fc2 GEMMs (comparison)
  if (fused_mlp < 0) {
    const char* e = getenv("DEVIT_FUSED_MLP");
    fused_mlp = (e && e[0] == '0') ? 0 : 1;
  }
  static int fused_proj_cache = kEnvUnread;  // DEVIT_FUSED_PROJ=0: separate proj GEMM (comparison)
  const bool fused_proj = env_int("DEVIT_FUSED_PROJ", 1, &fused_proj_cache) != 0;
  static int split_cache = kEnvUnread;  // DEVIT_SPLIT_RESID=0: fp32 residual stream between layers
  const bool split_resid = env_int("DEVIT_SPLIT_RESID", 1, &split_cache) != 0;
  bool resid_is_split = false;
  void* const qkv_ws = qkv;
  for (int l = 0; l < nl; ++l) {
    const devit_layer_desc& w = layers[l];
    const int hd = w.heads * 64;
    // an exported layer's QKV GEMM writes straight into the caller's buffer and the attention
    // kernel reads it from there: the export costs no copy
    qkv = (bufs.qkv_export && bufs.qkv_export[l]) ? bufs.qkv_export[l] : qkv_ws;
    // x -> LN1 -> y                                              (models/de_vit.py:113)
    if (!fold) {
      rc = devit_layernorm(x, w.ln1_g, w.ln1_b, y, M, D, ln_eps, opk, M * D, stream);
      if (rc) return rc;
    }
    // qkv = y Wqkv^T + b                                         (:67)
    base_gemm(&g, prec);
    g.m = static_cast<int>(M); g.n = 3 * hd;
    g.a = y; g.a_rows = static_cast<int>(M); g.a_cols = D; g.lda = D; g.a_plane_stride = M * D;
    g.b = w.w_qkv; g.b_rows = 3 * hd; g.b_cols = D; g.ldb = D;
    g.b_plane_stride = static_cast<long long>(3 * hd) * D;
    g.segs[0] = devit_gemm_seg{0, 0, 0, D};
    g.out = qkv; g.ldo = 3 * hd; g.out_kind = opk; g.out_plane_stride = M * 3 * hd;
    if ((rc = sync_debug("ln1", l, stream))) return rc;
    g.bias = w.b_qkv;
    if (fold) {
      g.ln_stats = stats; g.ln_parts = parts; g.ln_dim = D; g.ln_eps = ln_eps;
      g.ln_colsum = w.cs_qkv;
    }
    g.profile_tag = DEVIT_TAG_GEMM_QKV;
    rc = devit_gemm(&g, stream);
    if (rc) return rc;
    if ((rc = sync_debug("qkv gemm", l, stream))) return rc;
    // o = softmax(q k^T / 8) v  per kept head                    (:70-74)
    rc = devit_attention(prec, qkv, M * 3 * hd, o, M * hd, batch, tokens, w.heads, 0.125f,
                         stream);
    if (rc) return rc;
    if ((rc = sync_debug("attention", l, stream))) return rc;
    const int F = w.hidden_ld;
    const bool fuse_mlp = fold && fused_mlp && (D == 384 || D == 256);
    if (fuse_mlp && fused_proj) {
      // x1 = x + o Wproj^T + b ; x = x1 + gelu(LN2(x1) W1^T + b1) W2^T + b2 in ONE kernel: the
      // residual stream makes one fp32 round trip per layer, x1 / its bf16 copy / its row sums
      // stay on chip                                               (:81-82, :114, :35-47, :115)
      devit_mlp_args ma;
      std::memset(&ma, 0, sizeof(ma));
      ma.m = static_cast<int>(M); ma.dim = D; ma.hidden_ld = F;
      ma.w1 = w.w_fc1; ma.c1 = w.cs_fc1; ma.c2 = w.b_fc1; ma.ln_eps = ln_eps;
      ma.w2 = w.w_fc2; ma.b2 = w.b_fc2; ma.x = x;
      ma.o = o; ma.w_proj = w.w_proj; ma.b_proj = w.b_proj; ma.proj_k = hd;
      if (l + 1 < nl) { ma.xb_out = y; ma.stats_out = stats; }
      // between fused layers the residual stream travels as two bf16 planes (hi = y, lo in the
      // otherwise unused `hid` buffer): 4 instead of 6 bytes stored per element (an SM stores
      // ~32 B/clk); the first layer reads, and the last one writes, the fp32 stream
      if (split_resid) {
        if (resid_is_split) { ma.xb = y; ma.x_lo_in = hid; }
        if (l + 1 < nl) ma.x_lo_out = hid;
        resid_is_split = l + 1 < nl;
      }
      rc = devit_mlp_fused(&ma, stream);
      if (rc) return rc;
      parts = 4;  // the fused kernel emits one partial row sum per dim/4 columns
      if ((rc = sync_debug("fused proj + mlp", l, stream))) return rc;
      continue;
    }
    // x += o Wproj^T + b                                         (:81, :114)
    base_gemm(&g, prec);
    g.m = static_cast<int>(M); g.n = D;
    g.a = o; g.a_rows = static_cast<int>(M); g.a_cols = hd; g.lda = hd; g.a_plane_stride = M * hd;
    g.b = w.w_proj; g.b_rows = D; g.b_cols = hd; g.ldb = hd;
    g.b_plane_stride = static_cast<long long>(D) * hd;
    g.segs[0] = devit_gemm_seg{0, 0, 0, hd};
    g.out = x; g.ldo = D; g.out_kind = DEVIT_OUT_F32;
    g.bias = w.b_proj; g.resid = x; g.ldr = D;
    if (fold) {
      g.out_bf16 = y; g.ld_out_bf16 = D; g.stats_out = stats;
      parts = stat_parts;
    }
    g.profile_tag = DEVIT_TAG_GEMM_PROJ;
    rc = devit_gemm(&g, stream);
    if (rc) return rc;
    if ((rc = sync_debug("proj gemm", l, stream))) return rc;
    // x -> LN2 -> y                                              (:115)
    if (!fold) {
      rc = devit_layernorm(x, w.ln2_g, w.ln2_b, y, M, D, ln_eps, opk, M * D, stream);
      if (rc) return rc;
    }
    if (fuse_mlp) {
      // x += gelu(LN2(x) W1^T + b1) W2^T + b2 in one kernel, hidden kept on chip   (:35-47, :115)
      devit_mlp_args ma;
      std::memset(&ma, 0, sizeof(ma));
      ma.m = static_cast<int>(M); ma.dim = D; ma.hidden_ld = F;
      ma.xb = y; ma.w1 = w.w_fc1; ma.c1 = w.cs_fc1; ma.c2 = w.b_fc1;
      ma.ln_stats = stats; ma.ln_parts = parts; ma.ln_eps = ln_eps;
      ma.w2 = w.w_fc2; ma.b2 = w.b_fc2; ma.x = x;
      if (l + 1 < nl) { ma.xb_out = y; ma.stats_out = stats; }
      rc = devit_mlp_fused(&ma, stream);
      if (rc) return rc;
      parts = 4;  // the fused kernel emits one partial row sum per 96 columns
      if ((rc = sync_debug("fused mlp", l, stream))) return rc;
      continue;
    }
    // hid = gelu(y W1^T + b1), kept neurons only                 (:36-37)
    base_gemm(&g, prec);
    g.m = static_cast<int>(M); g.n = F;
    g.a = y; g.a_rows = static_cast<int>(M); g.a_cols = D; g.lda = D; g.a_plane_stride = M * D;
    g.b = w.w_fc1; g.b_rows = F; g.b_cols = D; g.ldb = D;
    g.b_plane_stride = static_cast<long long>(F) * D;
    g.segs[0] = devit_gemm_seg{0, 0, 0, D};
    g.out = hid; g.ldo = F; g.out_kind = opk; g.out_plane_stride = M * F;
    if ((rc = sync_debug("ln2", l, stream))) return rc;
    g.bias = w.b_fc1; g.act = DEVIT_ACT_GELU_ERF;
    if (fold) {
      g.ln_stats = stats; g.ln_parts = parts; g.ln_dim = D; g.ln_eps = ln_eps;
      g.ln_colsum = w.cs_fc1;
    }
    g.profile_tag = DEVIT_TAG_GEMM_FC1;
    rc = devit_gemm(&g, stream);
    if (rc) return rc;
    if ((rc = sync_debug("fc1 gemm", l, stream))) return rc;
    // x += hid W2^T + b2                                         (:45, :115)
    base_gemm(&g, prec);
    g.m = static_cast<int>(M); g.n = D;
    g.a = hid; g.a_rows = static_cast<int>(M); g.a_cols = F; g.lda = F; g.a_plane_stride = M * F;
    g.b = w.w_fc2; g.b_rows = D; g.b_cols = F; g.ldb = F;
    g.b_plane_stride = static_cast<long long>(D) * F;
    g.segs[0] = devit_gemm_seg{0, 0, 0, F};
    g.out = x; g.ldo = D; g.out_kind = DEVIT_OUT_F32;
    g.bias = w.b_fc2; g.resid = x; g.ldr = D;
    if (fold && l + 1 < nl) {  // the next layer's norm1 input (the final norm reads x itself)
      g.out_bf16 = y; g.ld_out_bf16 = D; g.stats_out = stats;
    }
    g.profile_tag = DEVIT_TAG_GEMM_FC2;
    rc = devit_gemm(&g, stream);
    if (rc) return rc;
    if ((rc = sync_debug("fc2 gemm", l, stream))) return rc;
  }
  return DEVIT_OK;
}

static int vit_forward_impl(const devit_vit_desc* d, const float* images, const void* patches,
                            int64_t patches_plane_stride, int32_t batch, void* workspace,
                            size_t workspace_bytes, float* feats_f32, void* feats_op,
                            int64_t feats_op_plane_stride, float* x_out, int32_t num_layers_run,
                            void* stream, const devit_vit_exports* exports = nullptr) {
  int rc = check_device();
  if (rc) return rc;
  VitLayout L{};
  rc = plan_layout(d, batch, &L);
  if (rc) return rc;
  DEVIT_REQUIRE((images || patches) && workspace, "devit_vit_forward: null pointer");
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 256 == 0,
                "devit_vit_forward: workspace must be 256-byte aligned");
  if (workspace_bytes < L.total)
    return set_error(DEVIT_ERR_WORKSPACE, "devit_vit_forward: workspace %zu < required %zu",
                     workspace_bytes, L.total);
  const int prec = d->precision;
  const int opk = prec == DEVIT_BF16 ? DEVIT_OUT_BF16 : DEVIT_OUT_F32_SPLIT;
  const int D = d->dim;
  const long long M = L.M;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(ws + L.off_x);
  void* y = ws + L.off_y;
  void* qkv = ws + L.off_qkv;
  void* o = ws + L.off_o;
  void* hid = ws + L.off_hid;
  const int kp = d->chans * 256;

  const int nl = (num_layers_run < 0 || num_layers_run > d->depth) ? d->depth : num_layers_run;
  int n_folded = 0;
  for (int l = 0; l < d->depth; ++l) n_folded += (d->layers[l].cs_qkv && d->layers[l].cs_fc1) ? 1 : 0;
  const bool fold = n_folded == d->depth;
  DEVIT_REQUIRE(n_folded == 0 || fold, "devit_vit_forward: cs_qkv / cs_fc1 must be set for all "
                "layers or for none");
  DEVIT_REQUIRE(!fold || (prec == DEVIT_BF16 && D % 128 == 0 && L.stat_parts <= 12),
                "devit_vit_forward: LayerNorm-folded weights need DEVIT_BF16 and dim a multiple of "
                "128 up to 768 (got dim %d)", D);
  float* stats = reinterpret_cast<float*>(ws + L.off_stats);
  int parts = 1;

  // ---- patch embedding as a residual GEMM over TOKEN rows (models/de_vit.py:258-264):
  //      A = token-row patch matrix (zero rows for cls / dist), x = pos (+ cls/dist - bias),
  //      x += A Wp^T + bias through the coalesced TMA epilogue.  With LayerNorm folding the same
  //      epilogue also emits the bf16 copy + row sums the first QKV GEMM consumes.
  const void* a_patch = patches;
  long long a_plane = patches_plane_stride;
  if (!a_patch) {
    a_plane = M * kp;
    rc = devit_im2col_tokens(images, qkv, batch, d->chans, d->img, d->num_prefix, opk, a_plane,
                             stream);
    if (rc) return rc;
    if ((rc = sync_debug("im2col", -1, stream))) return rc;
    a_patch = qkv;
  }
  // x = pos (+ cls/dist - bias): either materialised per image (devit_token_init) and added as a
  // plain residual, or -- when the host packed the wrapped [tokens + 31, D] table -- added by the
  // GEMM as a periodic residual straight from L2 (no 78 MB write + read of x at bs 256)
  if (!d->tok_table) {
    rc = devit_token_init(x, d->prefix, d->pos, d->b_patch, batch, L.tokens, D, d->num_prefix,
                          stream);
    if (rc) return rc;
    if ((rc = sync_debug("token init", -1, stream))) return rc;
  }
  devit_gemm_args g;
  base_gemm(&g, prec);
  g.m = static_cast<int>(M);
  g.n = D;
  g.a = a_patch; g.a_rows = static_cast<int>(M); g.a_cols = kp; g.lda = kp; g.a_plane_stride = a_plane;
  g.b = d->w_patch; g.b_rows = D; g.b_cols = kp; g.ldb = kp;
  g.b_plane_stride = static_cast<long long>(D) * kp;
  g.segs[0] = devit_gemm_seg{0, 0, 0, kp};
  g.out = x; g.ldo = D; g.out_kind = DEVIT_OUT_F32;
  g.bias = d->b_patch; g.resid = x; g.ldr = D;
  if (d->tok_table) { g.resid = d->tok_table; g.resid_period = L.tokens; }
  if (fold && nl > 0) {
    g.out_bf16 = y; g.ld_out_bf16 = D; g.stats_out = stats;
    parts = L.stat_parts;
  }
  g.profile_tag = DEVIT_TAG_GEMM_PATCH;
  rc = devit_gemm(&g, stream);
  if (rc) return rc;
  if ((rc = sync_debug("patch gemm", -1, stream))) return rc;

  // LayerNorm folding (bf16 mode, host packed gamma-folded weights): no LayerNorm kernel and no
  // normalised tensor at all.  `y` holds the bf16 copy of the residual stream, written by the
  // epilogue of whichever GEMM last updated x together with the rows' partial (sum, sum^2); the
  // QKV / fc1 GEMMs consume both and apply mean / rstd in their epilogue.
  {
    BlockBuffers bufs{x, y, qkv, o, hid, stats};
    if (exports) {
      DEVIT_REQUIRE(d->depth <= DEVIT_MAX_DEPTH, "devit_vit_forward_ex: depth %d > %d", d->depth,
                    DEVIT_MAX_DEPTH);
      for (int l = 0; l < nl; ++l)
        DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(exports->qkv[l]) % 256 == 0,
                      "devit_vit_forward_ex: qkv export buffers must be 256-byte aligned");
      bufs.qkv_export = exports->qkv;
    }
    rc = run_blocks(d->layers, nl, prec, D, d->ln_eps, batch, L.tokens, M, bufs, fold, parts,
                    L.stat_parts, stream);
    if (rc) return rc;
  }
  if (x_out) {
    DEVIT_CUDA_OK(cudaMemcpyAsync(x_out, x, static_cast<size_t>(M) * D * 4,
                                  cudaMemcpyDeviceToDevice,
                                  reinterpret_cast<cudaStream_t>(stream)));
  }
  // final norm on the returned rows only                         (:286-288)
  if (feats_f32 || feats_op) {
    rc = devit_gather_ln(x, d->norm_g, d->norm_b, feats_f32, feats_op, opk,
                         feats_op_plane_stride, batch, L.tokens, D, d->num_prefix, d->ln_eps,
                         exports ? exports->feats_kind_rows : 0, stream);
    if (rc) return rc;
  }
  return DEVIT_OK;
}

extern "C" int devit_vit_forward(const devit_vit_desc* d, const float* images, int32_t batch,
                                 void* workspace, size_t workspace_bytes, float* feats_f32,
                                 void* feats_op, int64_t feats_op_plane_stride, float* x_out,
                                 int32_t num_layers_run, void* stream) {
  DEVIT_REQUIRE(images, "devit_vit_forward: null images");
  return vit_forward_impl(d, images, nullptr, 0, batch, workspace, workspace_bytes, feats_f32,
                          feats_op, feats_op_plane_stride, x_out, num_layers_run, stream);
}

extern "C" int devit_vit_forward_patches(const devit_vit_desc* d, const void* patches,
                                         int64_t patches_plane_stride, int32_t batch,
                                         void* workspace, size_t workspace_bytes, float* feats_f32,
                                         void* feats_op, int64_t feats_op_plane_stride,
                                         float* x_out, int32_t num_layers_run, void* stream) {
  DEVIT_REQUIRE(patches, "devit_vit_forward_patches: null patches");
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(patches) % 16 == 0,
                "devit_vit_forward_patches: patches must be 16-byte aligned");
  return vit_forward_impl(d, nullptr, patches, patches_plane_stride, batch, workspace,
                          workspace_bytes, feats_f32, feats_op, feats_op_plane_stride, x_out,
                          num_layers_run, stream);
}

extern "C" int devit_vit_forward_ex(const devit_vit_desc* d, const float* images,
                                    const void* patches, int64_t patches_plane_stride,
                                    int32_t batch, void* workspace, size_t workspace_bytes,
                                    float* feats_f32, void* feats_op,
                                    int64_t feats_op_plane_stride, float* x_out,
                                    int32_t num_layers_run, const devit_vit_exports* exports,
                                    void* stream) {
  DEVIT_REQUIRE((images != nullptr) != (patches != nullptr),
                "devit_vit_forward_ex: pass exactly one of images / patches");
  DEVIT_REQUIRE(!patches || reinterpret_cast<uintptr_t>(patches) % 16 == 0,
                "devit_vit_forward_ex: patches must be 16-byte aligned");
  return vit_forward_impl(d, images, patches, patches_plane_stride, batch, workspace,
                          workspace_bytes, feats_f32, feats_op, feats_op_plane_stride, x_out,
                          num_layers_run, stream, exports);
}

// ------------------------------------------------------------------------------- CCT
namespace devit {

struct CctLayout {
  int tokens, max_hidden_ld, max_heads, stat_parts, planes, esz;
  long long M;
  size_t off_x, off_y, off_qkv, off_o, off_hid, off_stats, off_xn, off_a, off_c, off_p, total;
};

static int plan_cct(const devit_cct_desc* d, int batch, CctLayout* L) {
  DEVIT_REQUIRE(d && d->layers, "devit_cct: null descriptor");
  DEVIT_REQUIRE(d->precision == DEVIT_BF16 || d->precision == DEVIT_FP32,
                "devit_cct: bad precision %d", d->precision);
  DEVIT_REQUIRE(d->dim == 256 || d->dim == 384 || d->dim == 768,
                "devit_cct: dim %d not in {256,384,768}", d->dim);
  DEVIT_REQUIRE(d->n_conv >= 1 && d->n_conv <= 3 && d->depth > 0 && batch > 0 && d->chans > 0,
                "devit_cct: bad geometry");
  DEVIT_REQUIRE(d->img > 0 && d->img % (1 << d->n_conv) == 0, "devit_cct: image side %d", d->img);
  DEVIT_REQUIRE(d->conv_chans[d->n_conv - 1] == d->dim,
                "devit_cct: the last conv layer must produce dim channels");
  L->planes = d->precision == DEVIT_BF16 ? 1 : 2;
  L->esz = d->precision == DEVIT_BF16 ? 2 : 4;
  const size_t pe = static_cast<size_t>(L->planes) * L->esz;
  const int side = d->img >> d->n_conv;
  L->tokens = side * side;
  DEVIT_REQUIRE(L->tokens <= 256, "devit_cct: %d tokens exceed the attention kernel's 256", L->tokens);
  L->M = static_cast<long long>(batch) * L->tokens;
  L->max_heads = 0;
  L->max_hidden_ld = 0;
  for (int l = 0; l < d->depth; ++l) {
    const devit_layer_desc& y = d->layers[l];
    DEVIT_REQUIRE(y.heads >= 1 && y.heads * 64 <= d->dim, "devit_cct: layer %d heads %d", l, y.heads);
    DEVIT_REQUIRE(y.hidden >= 1 && y.hidden_ld >= y.hidden && y.hidden_ld % 16 == 0,
                  "devit_cct: layer %d hidden %d / ld %d", l, y.hidden, y.hidden_ld);
    if (y.heads > L->max_heads) L->max_heads = y.heads;
    if (y.hidden_ld > L->max_hidden_ld) L->max_hidden_ld = y.hidden_ld;
  }
  L->stat_parts = 2 * ((d->dim + 127) / 128);
  size_t a_b = 0, c_b = 0, p_b = 0;
  int hw = d->img, cin = d->chans;
  for (int i = 0; i < d->n_conv; ++i) {
    DEVIT_REQUIRE(d->w_conv[i] && d->conv_kpad[i] >= 9 * cin && d->conv_kpad[i] % 8 == 0 &&
                      d->conv_chans[i] % 4 == 0,
                  "devit_cct: conv layer %d (kpad %d for %d input channels)", i, d->conv_kpad[i], cin);
    const size_t rows = static_cast<size_t>(batch) * hw * hw;
    const size_t ab = rows * d->conv_kpad[i] * pe;
    const size_t cb = rows * d->conv_chans[i] * (d->precision == DEVIT_BF16 ? 2 : 4);
    const size_t pb = rows / 4 * d->conv_chans[i] * 4;
    if (ab > a_b) a_b = ab;
    if (cb > c_b) c_b = cb;
    if (i + 1 < d->n_conv && pb > p_b) p_b = pb;
    cin = d->conv_chans[i];
    hw >>= 1;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off = align_up(off + bytes, 256);
    return at;
  };
  L->off_x = take(static_cast<size_t>(L->M) * d->dim * 4);
  L->off_y = take(static_cast<size_t>(L->M) * d->dim * pe);
  L->off_qkv = take(static_cast<size_t>(L->M) * 3 * L->max_heads * 64 * pe);
  L->off_o = take(static_cast<size_t>(L->M) * L->max_heads * 64 * pe);
  L->off_hid = take(static_cast<size_t>(L->M) *
                    (L->max_hidden_ld > d->dim ? L->max_hidden_ld : d->dim) * pe);
  L->off_stats = take(static_cast<size_t>(L->M) * L->stat_parts * 2 * sizeof(float));
  L->off_xn = take(static_cast<size_t>(L->M) * d->dim * 4);
  L->off_a = take(a_b);
  L->off_c = take(c_b);
  L->off_p = take(p_b);
  L->total = off;
  return DEVIT_OK;
}

}  // namespace devit

extern "C" size_t devit_cct_workspace_bytes(const devit_cct_desc* desc, int32_t batch) {
  CctLayout L{};
  if (plan_cct(desc, batch, &L)) return 0;
  return L.total;
}

extern "C" int devit_cct_forward(const devit_cct_desc* d, const float* images, int32_t batch,
                                 void* workspace, size_t workspace_bytes, float* pooled,
                                 float* x_out, int32_t num_layers_run, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  CctLayout L{};
  rc = plan_cct(d, batch, &L);
  if (rc) return rc;
  DEVIT_REQUIRE(images && workspace, "devit_cct_forward: null pointer");
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 256 == 0,
                "devit_cct_forward: workspace must be 256-byte aligned");
  if (workspace_bytes < L.total)
    return set_error(DEVIT_ERR_WORKSPACE, "devit_cct_forward: workspace %zu < required %zu",
                     workspace_bytes, L.total);
  const int prec = d->precision;
  const int opk = prec == DEVIT_BF16 ? DEVIT_OUT_BF16 : DEVIT_OUT_F32_SPLIT;
  const int ck = prec == DEVIT_BF16 ? DEVIT_OUT_BF16 : DEVIT_OUT_F32;  // conv GEMM output
  const int D = d->dim;
  const long long M = L.M;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(ws + L.off_x);
  void* y = ws + L.off_y;
  float* stats = reinterpret_cast<float*>(ws + L.off_stats);
  float* xn = reinterpret_cast<float*>(ws + L.off_xn);
  void* sa = ws + L.off_a;
  void* sc = ws + L.off_c;
  float* sp = reinterpret_cast<float*>(ws + L.off_p);

  // ---- tokenizer: n_conv x [conv3x3 (im2col + GEMM, ReLU epilogue) -> max-pool 3/2/1]
  //      (models/utils/tokenizer.py:23-44); the last pool adds positional_emb and writes x.
  const float* in = images;
  int hw = d->img, cin = d->chans;
  long long sb = static_cast<long long>(cin) * hw * hw, scs = static_cast<long long>(hw) * hw,
            sy = hw, sx = 1;
  for (int i = 0; i < d->n_conv; ++i) {
    const long long rows = static_cast<long long>(batch) * hw * hw;
    const int kpad = d->conv_kpad[i], cout = d->conv_chans[i];
    rc = devit_im2col3x3(in, sa, batch, cin, hw, sb, scs, sy, sx, kpad, opk, rows * kpad, stream);
    if (rc) return rc;
    devit_gemm_args g;
    base_gemm(&g, prec);
    g.m = static_cast<int>(rows); g.n = cout;
    g.a = sa; g.a_rows = g.m; g.a_cols = kpad; g.lda = kpad; g.a_plane_stride = rows * kpad;
    g.b = d->w_conv[i]; g.b_rows = cout; g.b_cols = kpad; g.ldb = kpad;
    g.b_plane_stride = static_cast<long long>(cout) * kpad;
    g.segs[0] = devit_gemm_seg{0, 0, 0, kpad};
    g.out = sc; g.ldo = cout; g.out_kind = ck;
    g.act = DEVIT_ACT_RELU;
    g.profile_tag = DEVIT_TAG_GEMM_PATCH;
    rc = devit_gemm(&g, stream);
    if (rc) return rc;
    const bool last = i + 1 == d->n_conv;
    rc = devit_maxpool3x3s2_cl(sc, ck, last ? x : sp, last ? d->pos : nullptr, batch, hw, cout,
                               stream);
    if (rc) return rc;
    if ((rc = sync_debug("cct conv layer", i, stream))) return rc;
    hw >>= 1;
    cin = cout;
    in = sp;  // channels-last [B, hw, hw, cout]
    sb = static_cast<long long>(hw) * hw * cout; scs = 1; sy = static_cast<long long>(hw) * cout; sx = cout;
  }

  // ---- transformer blocks (shared with the ViT path)
  const int nl = (num_layers_run < 0 || num_layers_run > d->depth) ? d->depth : num_layers_run;
  int n_folded = 0;
  for (int l = 0; l < d->depth; ++l) n_folded += (d->layers[l].cs_qkv && d->layers[l].cs_fc1) ? 1 : 0;
  const bool fold = n_folded == d->depth;
  DEVIT_REQUIRE(n_folded == 0 || fold, "devit_cct_forward: cs_qkv / cs_fc1 must be set for all "
                "layers or for none");
  DEVIT_REQUIRE(!fold || (prec == DEVIT_BF16 && D % 128 == 0 && L.stat_parts <= 12),
                "devit_cct_forward: LayerNorm-folded weights need DEVIT_BF16 and dim %% 128 == 0");
  if (fold && nl > 0) {
    rc = devit_rowstats(x, y, stats, M, D, stream);
    if (rc) return rc;
  }
  {
    BlockBuffers bufs{x, y, ws + L.off_qkv, ws + L.off_o, ws + L.off_hid, stats};
    rc = run_blocks(d->layers, nl, prec, D, d->ln_eps, batch, L.tokens, M, bufs, fold, 1,
                    L.stat_parts, stream);
    if (rc) return rc;
  }
  if (x_out)
    DEVIT_CUDA_OK(cudaMemcpyAsync(x_out, x, static_cast<size_t>(M) * D * 4,
                                  cudaMemcpyDeviceToDevice,
                                  reinterpret_cast<cudaStream_t>(stream)));
  // ---- final norm over every token + sequence pooling (models/utils/transformers.py:470-475)
  if (pooled) {
    rc = devit_layernorm(x, d->norm_g, d->norm_b, xn, M, D, d->ln_eps, DEVIT_OUT_F32, 0, stream);
    if (rc) return rc;
    rc = devit_seqpool(xn, d->pool_w, d->pool_b, pooled, batch, L.tokens, D, stream);
    if (rc) return rc;
  }
  return DEVIT_OK;
}
