// Memory-bound row kernels: LayerNorm (one warp per row, 128-bit loads, shuffle reductions),
// the final-norm + cls/dist gather, cls/dist token rows, and the patch im2col + cast.
// All are pure streaming kernels; their roofline is HBM bandwidth (DESIGN.md gives bytes/row).
#include "common.cuh"
#include "ptx.cuh"

namespace devit {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Normalises one row held as V float4 per lane (dim = V*128) and writes it in `out_kind`.
template <int V>
__device__ __forceinline__ void ln_row(const float* __restrict__ xr, const float* __restrict__ g,
                                       const float* __restrict__ b, float eps, int lane,
                                       int out_kind, void* y_row, long long plane,
                                       float* y_f32_row) {
  constexpr int D = V * 128;
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = *(reinterpret_cast<const float4*>(xr) + lane + 32 * i);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int e4 = lane + 32 * i;
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + e4);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + e4);
    float4 o;
    o.x = v[i].x * rstd * gg.x + bb.x;
    o.y = v[i].y * rstd * gg.y + bb.y;
    o.z = v[i].z * rstd * gg.z + bb.z;
    o.w = v[i].w * rstd * gg.w + bb.w;
    if (y_f32_row) *(reinterpret_cast<float4*>(y_f32_row) + e4) = o;
    if (y_row) {
      if (out_kind == DEVIT_OUT_BF16) {
        uint2 t;
        t.x = pack_bf16x2(o.x, o.y);
        t.y = pack_bf16x2(o.z, o.w);
        *(reinterpret_cast<uint2*>(y_row) + e4) = t;
      } else if (out_kind == DEVIT_OUT_F32) {
        *(reinterpret_cast<float4*>(y_row) + e4) = o;
      } else {
        float4 h = make_float4(tf32_hi(o.x), tf32_hi(o.y), tf32_hi(o.z), tf32_hi(o.w));
        *(reinterpret_cast<float4*>(y_row) + e4) = h;
        *(reinterpret_cast<float4*>(static_cast<float*>(y_row) + plane) + e4) =
            make_float4(o.x - h.x, o.y - h.y, o.z - h.z, o.w - h.w);
      }
    }
  }
}

template <int V>
__global__ void __launch_bounds__(256)
ln_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
          void* __restrict__ y, long long rows, float eps, int out_kind, long long plane) {
  constexpr int D = V * 128;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int esz = out_kind == DEVIT_OUT_BF16 ? 2 : 4;
  ln_row<V>(x + row * D, g, b, eps, threadIdx.x & 31, out_kind,
            static_cast<uint8_t*>(y) + row * D * esz, plane, nullptr);
}

// bf16 copy of the fp32 residual stream + (sum, sum of squares) per row: the one-part input of
// a LayerNorm-folded GEMM (include/devit_b200.h, devit_rowstats).
template <int V>
__global__ void __launch_bounds__(256)
rowstats_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                float* __restrict__ stats, long long rows) {
  constexpr int D = V * 128;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  uint2* br = reinterpret_cast<uint2*>(xb + row * D);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 v = xr[lane + 32 * i];
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    uint2 t;
    t.x = pack_bf16x2(v.x, v.y);
    t.y = pack_bf16x2(v.z, v.w);
    br[lane + 32 * i] = t;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) *reinterpret_cast<float2*>(stats + 2 * row) = make_float2(s1, s2);
}

template <int V>
__global__ void __launch_bounds__(256)
gather_ln_kernel(const float* __restrict__ x, const float* __restrict__ g,
                 const float* __restrict__ b, float* __restrict__ feats_f32,
                 void* __restrict__ feats_op, int out_kind, long long plane, int batch, int tokens,
                 int num_prefix, float eps, int kind_rows) {
  constexpr int D = V * 128;
  const int idx = blockIdx.x * 8 + (threadIdx.x >> 5);  // = img * num_prefix + j
  if (idx >= batch * num_prefix) return;
  const int img = idx / num_prefix, j = idx - img * num_prefix;
  const long long orow = static_cast<long long>(j) * kind_rows + img;
  const int esz = out_kind == DEVIT_OUT_BF16 ? 2 : 4;
  ln_row<V>(x + (static_cast<long long>(img) * tokens + j) * D, g, b, eps, threadIdx.x & 31,
            out_kind, feats_op ? static_cast<uint8_t*>(feats_op) + orow * D * esz : nullptr, plane,
            feats_f32 ? feats_f32 + orow * D : nullptr);
}

__global__ void token_prefix_kernel(float* __restrict__ x, const float* __restrict__ prefix,
                                    const float* __restrict__ pos, int batch, int tokens, int dim,
                                    int num_prefix) {
  const int per_img = num_prefix * dim;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(batch) * per_img) return;
  const int img = static_cast<int>(i / per_img);
  const int r = static_cast<int>(i - static_cast<long long>(img) * per_img);
  x[static_cast<long long>(img) * tokens * dim + r] = prefix[r] + pos[r];
}

// x[b, j, :] = pos[j, :] + (j < num_prefix ? prefix[j, :] - bias : 0): the residual-stream
// initial value the patch-embedding GEMM (+ bias) is accumulated onto (include/devit_b200.h).
__global__ void __launch_bounds__(256)
token_init_kernel(float* __restrict__ x, const float* __restrict__ prefix,
                  const float* __restrict__ pos, const float* __restrict__ bias, int batch,
                  int tokens, int dim4, int num_prefix) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_img = static_cast<long long>(tokens) * dim4;
  if (i >= batch * per_img) return;
  const int r = static_cast<int>(i % per_img);  // (token, float4 column) inside the image
  const int j = r / dim4, c = r - j * dim4;
  float4 v = __ldg(reinterpret_cast<const float4*>(pos) + r);
  if (j < num_prefix) {
    const float4 pf = __ldg(reinterpret_cast<const float4*>(prefix) + r);
    v.x += pf.x; v.y += pf.y; v.z += pf.z; v.w += pf.w;
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c);
      v.x -= b.x; v.y -= b.y; v.z -= b.z; v.w -= b.w;
    }
  }
  reinterpret_cast<float4*>(x)[i] = v;
}

// One thread converts 8 horizontally adjacent pixels of one patch row: two independent float4
// loads (32 B, one full sector: enough bytes in flight to cover the HBM latency at 2048
// threads/SM) and one 16-byte store (bf16).  Threads are numbered in OUTPUT order
// (image, patch, channel, py, half row), so consecutive threads write consecutive 16-byte chunks
// of the patch matrix: every warp stores 512 contiguous bytes and reads 16 x 64 B row pieces.
__global__ void __launch_bounds__(256)
im2col16_kernel(const float* __restrict__ img, void* __restrict__ a, int batch, int chans, int hw,
                int out_kind, long long plane, int row_off, int rows_per_img) {
  const int g = hw >> 4;
  const int P = g * g;
  const long long total = static_cast<long long>(batch) * P * chans * 32;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int hx = static_cast<int>(i & 1);
  const int py = static_cast<int>(i >> 1) & 15;
  long long t = i >> 5;
  const int c = static_cast<int>(t % chans);
  t /= chans;
  const int pidx = static_cast<int>(t % P);
  const int b = static_cast<int>(t / P);
  const int gy = pidx / g, gx = pidx - gy * g;
  const float4* src = reinterpret_cast<const float4*>(
      img + ((static_cast<long long>(b) * chans + c) * hw + gy * 16 + py) * hw + gx * 16 + hx * 8);
  const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
  const long long m = static_cast<long long>(b) * rows_per_img + row_off + pidx;
  const long long o = m * (static_cast<long long>(chans) * 256) + c * 256 + py * 16 + hx * 8;
  if (out_kind == DEVIT_OUT_BF16) {
    uint4 pk;
    pk.x = pack_bf16x2(v0.x, v0.y);
    pk.y = pack_bf16x2(v0.z, v0.w);
    pk.z = pack_bf16x2(v1.x, v1.y);
    pk.w = pack_bf16x2(v1.z, v1.w);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a) + o) = pk;
  } else if (out_kind == DEVIT_OUT_F32) {
    float4* d = reinterpret_cast<float4*>(static_cast<float*>(a) + o);
    d[0] = v0;
    d[1] = v1;
  } else {
    const float4 h0 = make_float4(tf32_hi(v0.x), tf32_hi(v0.y), tf32_hi(v0.z), tf32_hi(v0.w));
    const float4 h1 = make_float4(tf32_hi(v1.x), tf32_hi(v1.y), tf32_hi(v1.z), tf32_hi(v1.w));
    float4* dh = reinterpret_cast<float4*>(static_cast<float*>(a) + o);
    float4* dl = reinterpret_cast<float4*>(static_cast<float*>(a) + o + plane);
    dh[0] = h0;
    dh[1] = h1;
    dl[0] = make_float4(v0.x - h0.x, v0.y - h0.y, v0.z - h0.z, v0.w - h0.w);
    dl[1] = make_float4(v1.x - h1.x, v1.y - h1.y, v1.z - h1.z, v1.w - h1.w);
  }
}

}  // namespace devit

using namespace devit;

extern "C" int devit_layernorm(const float* x, const float* gamma, const float* beta, void* y,
                               int64_t rows, int32_t dim, float eps, int32_t out_kind,
                               int64_t out_plane_stride, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(x && gamma && beta && y, "devit_layernorm: null pointer");
  DEVIT_REQUIRE(rows > 0, "devit_layernorm: rows must be > 0");
  DEVIT_REQUIRE(out_kind >= 0 && out_kind <= 2, "devit_layernorm: bad out_kind %d", out_kind);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  ProfScope ps(kTagLayerNorm, stream);
  switch (dim) {
    case 256: ln_kernel<2><<<grid, 256, 0, stream>>>(x, gamma, beta, y, rows, eps, out_kind, out_plane_stride); break;
    case 384: ln_kernel<3><<<grid, 256, 0, stream>>>(x, gamma, beta, y, rows, eps, out_kind, out_plane_stride); break;
    case 768: ln_kernel<6><<<grid, 256, 0, stream>>>(x, gamma, beta, y, rows, eps, out_kind, out_plane_stride); break;
    default: return set_error(DEVIT_ERR_ARG, "devit_layernorm: dim %d not in {256,384,768}", dim);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" int devit_rowstats(const float* x, void* xb, float* stats, int64_t rows, int32_t dim,
                              void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(x && xb && stats, "devit_rowstats: null pointer");
  DEVIT_REQUIRE(rows > 0, "devit_rowstats: rows must be > 0");
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  __nv_bfloat16* b = static_cast<__nv_bfloat16*>(xb);
  ProfScope ps(kTagLayerNorm, stream);
  switch (dim) {
    case 256: rowstats_kernel<2><<<grid, 256, 0, stream>>>(x, b, stats, rows); break;
    case 384: rowstats_kernel<3><<<grid, 256, 0, stream>>>(x, b, stats, rows); break;
    case 768: rowstats_kernel<6><<<grid, 256, 0, stream>>>(x, b, stats, rows); break;
    default: return set_error(DEVIT_ERR_ARG, "devit_rowstats: dim %d not in {256,384,768}", dim);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" int devit_gather_ln(const float* x, const float* gamma, const float* beta,
                               float* feats_f32, void* feats_op, int32_t out_kind,
                               int64_t out_plane_stride, int32_t batch, int32_t tokens,
                               int32_t dim, int32_t num_prefix, float eps, int32_t kind_rows,
                               void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(x && gamma && beta && (feats_f32 || feats_op), "devit_gather_ln: null pointer");
  DEVIT_REQUIRE(batch > 0 && num_prefix > 0 && num_prefix <= tokens, "devit_gather_ln: bad shape");
  if (kind_rows <= 0) kind_rows = batch;
  DEVIT_REQUIRE(kind_rows >= batch, "devit_gather_ln: kind_rows %d < batch %d", kind_rows, batch);
  const unsigned grid = static_cast<unsigned>((batch * num_prefix + 7) / 8);
  ProfScope ps(kTagGatherLn, stream);
  switch (dim) {
    case 256: gather_ln_kernel<2><<<grid, 256, 0, stream>>>(x, gamma, beta, feats_f32, feats_op, out_kind, out_plane_stride, batch, tokens, num_prefix, eps, kind_rows); break;
    case 384: gather_ln_kernel<3><<<grid, 256, 0, stream>>>(x, gamma, beta, feats_f32, feats_op, out_kind, out_plane_stride, batch, tokens, num_prefix, eps, kind_rows); break;
    case 768: gather_ln_kernel<6><<<grid, 256, 0, stream>>>(x, gamma, beta, feats_f32, feats_op, out_kind, out_plane_stride, batch, tokens, num_prefix, eps, kind_rows); break;
    default: return set_error(DEVIT_ERR_ARG, "devit_gather_ln: dim %d not in {256,384,768}", dim);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" int devit_token_prefix(float* x, const float* prefix, const float* pos, int32_t batch,
                                  int32_t tokens, int32_t dim, int32_t num_prefix,
                                  void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(x && prefix && pos, "devit_token_prefix: null pointer");
  DEVIT_REQUIRE(batch > 0 && num_prefix > 0 && num_prefix <= tokens && dim > 0,
                "devit_token_prefix: bad shape");
  const long long total = static_cast<long long>(batch) * num_prefix * dim;
  {
    ProfScope ps(kTagPrefix, stream);
    token_prefix_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        x, prefix, pos, batch, tokens, dim, num_prefix);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" int devit_im2col_patch16(const float* images, void* a, int32_t batch, int32_t chans,
                                    int32_t hw, int32_t out_kind, int64_t out_plane_stride,
                                    void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(images && a, "devit_im2col_patch16: null pointer");
  DEVIT_REQUIRE(batch > 0 && chans > 0 && hw > 0 && hw % 16 == 0,
                "devit_im2col_patch16: image side %d must be a positive multiple of 16", hw);
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(images) % 16 == 0,
                "devit_im2col_patch16: images must be 16-byte aligned");
  const long long total = static_cast<long long>(batch) * chans * hw * (hw / 8);
  {
    ProfScope ps(kTagIm2col, stream);
    im2col16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        images, a, batch, chans, hw, out_kind, out_plane_stride, 0, (hw / 16) * (hw / 16));
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" int devit_im2col_tokens(const float* images, void* a, int32_t batch, int32_t chans,
                                   int32_t hw, int32_t num_prefix, int32_t out_kind,
                                   int64_t out_plane_stride, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(images && a, "devit_im2col_tokens: null pointer");
  DEVIT_REQUIRE(batch > 0 && chans > 0 && hw > 0 && hw % 16 == 0 && num_prefix >= 0,
                "devit_im2col_tokens: image side %d must be a positive multiple of 16", hw);
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(images) % 16 == 0,
                "devit_im2col_tokens: images must be 16-byte aligned");
  const int tokens = num_prefix + (hw / 16) * (hw / 16);
  const size_t esz = out_kind == DEVIT_OUT_BF16 ? 2 : 4;
  const size_t k = static_cast<size_t>(chans) * 256;
  const long long total = static_cast<long long>(batch) * chans * hw * (hw / 8);
  {
    ProfScope ps(kTagIm2col, stream);
    if (num_prefix > 0) {  // zero rows for the cls / dist tokens of every image
      DEVIT_CUDA_OK(cudaMemset2DAsync(a, tokens * k * esz, 0, num_prefix * k * esz, batch, stream));
      if (out_kind == DEVIT_OUT_F32_SPLIT)
        DEVIT_CUDA_OK(cudaMemset2DAsync(static_cast<float*>(a) + out_plane_stride, tokens * k * esz,
                                        0, num_prefix * k * esz, batch, stream));
    }
    im2col16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        images, a, batch, chans, hw, out_kind, out_plane_stride, num_prefix, tokens);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" int devit_token_init(float* x, const float* prefix, const float* pos,
                                const float* bias, int32_t batch, int32_t tokens, int32_t dim,
                                int32_t num_prefix, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(x && prefix && pos, "devit_token_init: null pointer");
  DEVIT_REQUIRE(batch > 0 && num_prefix >= 0 && num_prefix <= tokens && dim > 0 && dim % 4 == 0,
                "devit_token_init: bad shape");
  const long long total = static_cast<long long>(batch) * tokens * (dim / 4);
  {
    ProfScope ps(kTagPrefix, stream);
    token_init_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        x, prefix, pos, bias, batch, tokens, dim / 4, num_prefix);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}
