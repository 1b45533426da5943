// devit_pack_layer: fp32 master parameters of one transformer Block + the head / neuron gates
// -> the gate-compacted, LayerNorm-folded GEMM operands a devit_layer_desc points at
// (include/devit_b200.h).  This is the C-ABI form of what devit_b200/packing.py does with torch
// ops, so that a host that is not Python can build valid descriptors.
//
// The reference computes every head / neuron densely and multiplies by a 0/1 gate
// (models/de_vit.py:41-43, :77-79); dropping a gated unit is numerically identical, so the pack
// keeps only
//   * the kept heads' q/k/v rows of qkv.weight / bias and their columns of proj.weight (scaled by
//     the gate value), in ascending head order (core/imp_rank.py:132-153 decides WHICH heads);
//   * the kept neurons' rows of fc1.weight / bias and their columns of fc2.weight (scaled by the
//     gate value), zero-padded to a multiple of 16 (core/imp_rank.py:50-71).
// LayerNorm folding (devit_gemm_args.ln_stats): W' = W .* gamma (per input column),
// c1[n] = sum_k bf16(W'[n,k]) (what the tensor core will see), c2[n] = b[n] + sum_k W[n,k] beta[k].
// One-off work when a gate or a parameter changes; not on the per-batch path.
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace devit {

static inline size_t pack_align(size_t v) { return (v + 255) / 256 * 256; }

// out[r, :] = W[row_idx[r], :] .* gamma (row_idx[r] < 0: zero row), bf16 or hi/lo split;
// c1[r] = sum_k bf16(out[r, k]), c2[r] = bias[row] + sum_k W[row, k] beta[k]   (c1 / c2 optional)
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ w, int ld, int cols, const int* __restrict__ row_idx,
                 int rows_out, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ bias, int out_split, void* __restrict__ out,
                 long long plane, float* __restrict__ c1, float* __restrict__ c2) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows_out) return;
  const int sr = row_idx[r];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < cols; c += 32) {
    float raw = 0.f, v = 0.f;
    if (sr >= 0) {
      raw = w[static_cast<long long>(sr) * ld + c];
      v = gamma ? raw * gamma[c] : raw;
    }
    const long long o = static_cast<long long>(r) * cols + c;
    if (out_split) {
      const float hi = tf32_hi(v);
      static_cast<float*>(out)[o] = hi;
      static_cast<float*>(out)[o + plane] = v - hi;
    } else {
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      static_cast<__nv_bfloat16*>(out)[o] = b;
      s1 += __bfloat162float(b);
    }
    if (beta) s2 = fmaf(raw, beta[c], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    if (c1) c1[r] = s1;
    if (c2) c2[r] = (sr >= 0 && bias ? bias[sr] : 0.f) + (sr >= 0 ? s2 : 0.f);
  }
}

// out[r, c] = W[r, col_idx[c]] * col_scale[c]  (col_idx[c] < 0: zero column), r < rows
__global__ void __launch_bounds__(256)
pack_cols_kernel(const float* __restrict__ w, int ld, int rows, const int* __restrict__ col_idx,
                 const float* __restrict__ col_scale, int cols_out, int out_split,
                 void* __restrict__ out, long long plane) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * cols_out) return;
  const int r = static_cast<int>(i / cols_out), c = static_cast<int>(i - static_cast<long long>(r) * cols_out);
  const int sc = col_idx[c];
  const float v = sc >= 0 ? w[static_cast<long long>(r) * ld + sc] * col_scale[c] : 0.f;
  if (out_split) {
    const float hi = tf32_hi(v);
    static_cast<float*>(out)[i] = hi;
    static_cast<float*>(out)[i + plane] = v - hi;
  } else {
    static_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  }
}

struct PackPlan {
  std::vector<int> heads, neurons;        // kept indices, ascending
  std::vector<float> head_scale, neuron_scale;
  int hd = 0, f = 0, f_ld = 0;
  size_t esz = 2;                         // bytes per operand element (both planes together)
  size_t off_idx_qkv, off_idx_proj, off_scale_proj, off_idx_fc1, off_idx_fc2, off_scale_fc2;
  size_t off_wqkv, off_bqkv, off_cqkv, off_wproj, off_wfc1, off_bfc1, off_cfc1, off_wfc2, total;
};

static int make_plan(const devit_block_weights* w, int precision, PackPlan* P) {
  DEVIT_REQUIRE(w, "devit_pack_layer: null weights");
  DEVIT_REQUIRE(precision == DEVIT_BF16 || precision == DEVIT_FP32,
                "devit_pack_layer: bad precision %d", precision);
  DEVIT_REQUIRE(w->dim > 0 && w->num_heads > 0 && w->hidden > 0 && w->dim % w->num_heads == 0,
                "devit_pack_layer: bad geometry");
  DEVIT_REQUIRE(w->dim / w->num_heads == 64, "devit_pack_layer: head_dim %d unsupported (64 only)",
                w->dim / w->num_heads);
  P->heads.clear(); P->neurons.clear(); P->head_scale.clear(); P->neuron_scale.clear();
  for (int h = 0; h < w->num_heads; ++h) {
    const float g = w->head_gate ? w->head_gate[h] : 1.f;
    if (g != 0.f) { P->heads.push_back(h); P->head_scale.push_back(g); }
  }
  if (P->heads.empty()) {  // every head gated off: keep one, with zeroed proj columns
    P->heads.push_back(0);
    P->head_scale.push_back(0.f);
  }
  for (int n = 0; n < w->hidden; ++n) {
    const float g = w->neuron_gate ? w->neuron_gate[n] : 1.f;
    if (g != 0.f) { P->neurons.push_back(n); P->neuron_scale.push_back(g); }
  }
  P->hd = static_cast<int>(P->heads.size()) * 64;
  P->f = static_cast<int>(P->neurons.size());
  P->f_ld = P->f < 16 ? 16 : (P->f + 15) / 16 * 16;
  P->esz = precision == DEVIT_BF16 ? 2 : 8;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off = pack_align(off + bytes); return at; };
  const size_t D = w->dim;
  P->off_idx_qkv = take(3 * P->hd * sizeof(int));
  P->off_idx_proj = take(P->hd * sizeof(int));
  P->off_scale_proj = take(P->hd * sizeof(float));
  P->off_idx_fc1 = take(P->f_ld * sizeof(int));
  P->off_idx_fc2 = take(P->f_ld * sizeof(int));
  P->off_scale_fc2 = take(P->f_ld * sizeof(float));
  P->off_wqkv = take(3 * P->hd * D * P->esz);
  P->off_bqkv = take(3 * P->hd * sizeof(float));
  P->off_cqkv = take(3 * P->hd * sizeof(float));
  P->off_wproj = take(D * P->hd * P->esz);
  P->off_wfc1 = take(static_cast<size_t>(P->f_ld) * D * P->esz);
  P->off_bfc1 = take(P->f_ld * sizeof(float));
  P->off_cfc1 = take(P->f_ld * sizeof(float));
  P->off_wfc2 = take(D * static_cast<size_t>(P->f_ld) * P->esz);
  P->total = off;
  return DEVIT_OK;
}

}  // namespace devit

using namespace devit;

extern "C" size_t devit_pack_layer_bytes(const devit_block_weights* w, int32_t precision) {
  PackPlan P;
  if (make_plan(w, precision, &P)) return 0;
  return P.total;
}

extern "C" int devit_pack_layer(const devit_block_weights* w, int32_t precision, int32_t fold_ln,
                                void* packed, size_t packed_bytes, devit_layer_desc* out,
                                int32_t* kept_heads, int32_t* num_kept_heads,
                                int32_t* kept_neurons, int32_t* num_kept_neurons, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  PackPlan P;
  int rc = make_plan(w, precision, &P);
  if (rc) return rc;
  DEVIT_REQUIRE(packed && out, "devit_pack_layer: null output");
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(packed) % 256 == 0,
                "devit_pack_layer: packed buffer must be 256-byte aligned");
  if (packed_bytes < P.total)
    return set_error(DEVIT_ERR_WORKSPACE, "devit_pack_layer: buffer %zu < required %zu",
                     packed_bytes, P.total);
  DEVIT_REQUIRE(w->ln1_g && w->ln1_b && w->w_qkv && w->w_proj && w->b_proj && w->ln2_g &&
                    w->ln2_b && w->w_fc1 && w->b_fc1 && w->w_fc2 && w->b_fc2,
                "devit_pack_layer: null parameter pointer");
  const bool fold = fold_ln != 0;
  DEVIT_REQUIRE(!fold || precision == DEVIT_BF16,
                "devit_pack_layer: LayerNorm folding needs DEVIT_BF16 operands");
  const int D = w->dim, hd = P.hd, f_ld = P.f_ld;
  const int split = precision == DEVIT_FP32 ? 1 : 0;
  uint8_t* base = static_cast<uint8_t*>(packed);

  // ---- index / scale tables (host -> device; the copies are complete when the call returns)
  std::vector<int> idx_qkv(3 * hd), idx_proj(hd), idx_fc1(f_ld, -1), idx_fc2(f_ld, -1);
  std::vector<float> scale_proj(hd), scale_fc2(f_ld, 0.f);
  for (size_t i = 0; i < P.heads.size(); ++i)
    for (int d = 0; d < 64; ++d) {
      const int col = P.heads[i] * 64 + d;
      idx_proj[i * 64 + d] = col;
      scale_proj[i * 64 + d] = P.head_scale[i];
      for (int which = 0; which < 3; ++which) idx_qkv[which * hd + i * 64 + d] = which * D + col;
    }
  for (int i = 0; i < P.f; ++i) {
    idx_fc1[i] = idx_fc2[i] = P.neurons[i];
    scale_fc2[i] = P.neuron_scale[i];
  }
  auto up = [&](size_t off, const void* src, size_t bytes) {
    return cudaMemcpyAsync(base + off, src, bytes, cudaMemcpyHostToDevice, stream);
  };
  DEVIT_CUDA_OK(up(P.off_idx_qkv, idx_qkv.data(), idx_qkv.size() * sizeof(int)));
  DEVIT_CUDA_OK(up(P.off_idx_proj, idx_proj.data(), idx_proj.size() * sizeof(int)));
  DEVIT_CUDA_OK(up(P.off_scale_proj, scale_proj.data(), scale_proj.size() * sizeof(float)));
  DEVIT_CUDA_OK(up(P.off_idx_fc1, idx_fc1.data(), idx_fc1.size() * sizeof(int)));
  DEVIT_CUDA_OK(up(P.off_idx_fc2, idx_fc2.data(), idx_fc2.size() * sizeof(int)));
  DEVIT_CUDA_OK(up(P.off_scale_fc2, scale_fc2.data(), scale_fc2.size() * sizeof(float)));
  const int* d_idx_qkv = reinterpret_cast<const int*>(base + P.off_idx_qkv);
  const int* d_idx_proj = reinterpret_cast<const int*>(base + P.off_idx_proj);
  const float* d_scale_proj = reinterpret_cast<const float*>(base + P.off_scale_proj);
  const int* d_idx_fc1 = reinterpret_cast<const int*>(base + P.off_idx_fc1);
  const int* d_idx_fc2 = reinterpret_cast<const int*>(base + P.off_idx_fc2);
  const float* d_scale_fc2 = reinterpret_cast<const float*>(base + P.off_scale_fc2);
  float* b_qkv = reinterpret_cast<float*>(base + P.off_bqkv);
  float* c_qkv = reinterpret_cast<float*>(base + P.off_cqkv);
  float* b_fc1 = reinterpret_cast<float*>(base + P.off_bfc1);
  float* c_fc1 = reinterpret_cast<float*>(base + P.off_cfc1);

  // ---- qkv rows / fc1 rows (+ fold), proj columns / fc2 columns
  pack_rows_kernel<<<(3 * hd + 7) / 8, 256, 0, stream>>>(
      w->w_qkv, D, D, d_idx_qkv, 3 * hd, fold ? w->ln1_g : nullptr, fold ? w->ln1_b : nullptr,
      w->b_qkv, split, base + P.off_wqkv, static_cast<long long>(3 * hd) * D,
      fold ? c_qkv : nullptr, b_qkv);
  pack_rows_kernel<<<(f_ld + 7) / 8, 256, 0, stream>>>(
      w->w_fc1, D, D, d_idx_fc1, f_ld, fold ? w->ln2_g : nullptr, fold ? w->ln2_b : nullptr,
      w->b_fc1, split, base + P.off_wfc1, static_cast<long long>(f_ld) * D,
      fold ? c_fc1 : nullptr, b_fc1);
  {
    const long long n = static_cast<long long>(D) * hd;
    pack_cols_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        w->w_proj, D, D, d_idx_proj, d_scale_proj, hd, split, base + P.off_wproj, n);
  }
  {
    const long long n = static_cast<long long>(D) * f_ld;
    pack_cols_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        w->w_fc2, w->hidden, D, d_idx_fc2, d_scale_fc2, f_ld, split, base + P.off_wfc2, n);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch(4);
  DEVIT_CUDA_OK(cudaStreamSynchronize(stream));  // host tables go out of scope

  std::memset(out, 0, sizeof(*out));
  out->heads = static_cast<int>(P.heads.size());
  out->hidden = P.f > 0 ? P.f : 1;
  out->hidden_ld = f_ld;
  out->ln1_g = w->ln1_g; out->ln1_b = w->ln1_b;
  out->ln2_g = w->ln2_g; out->ln2_b = w->ln2_b;
  out->w_qkv = base + P.off_wqkv; out->b_qkv = b_qkv;
  out->w_proj = base + P.off_wproj; out->b_proj = w->b_proj;
  out->w_fc1 = base + P.off_wfc1; out->b_fc1 = b_fc1;
  out->w_fc2 = base + P.off_wfc2; out->b_fc2 = w->b_fc2;
  out->cs_qkv = fold ? c_qkv : nullptr;
  out->cs_fc1 = fold ? c_fc1 : nullptr;
  if (num_kept_heads) *num_kept_heads = static_cast<int>(P.heads.size());
  if (kept_heads) std::memcpy(kept_heads, P.heads.data(), P.heads.size() * sizeof(int));
  if (num_kept_neurons) *num_kept_neurons = P.f;
  if (kept_neurons && P.f) std::memcpy(kept_neurons, P.neurons.data(), P.f * sizeof(int));
  return DEVIT_OK;
}
