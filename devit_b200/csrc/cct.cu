// CCT (Compact Convolutional Transformer) front end and head, see include/devit_b200.h
// (devit_cct_forward): the convolutional tokenizer as im2col + tcgen05 GEMM (ReLU epilogue) +
// channels-last max-pool, and the sequence-pooling head.  The transformer blocks in between
// are the shared run of csrc/forward.cu.  All kernels here are memory-bound streaming kernels.
#include "common.cuh"
#include "ptx.cuh"

namespace devit {

// A[m, k] for the 3x3 / stride 1 / pad 1 convolution: m = (b, y, x), k = (ky*3 + kx)*C + c,
// zero beyond 9*C (K padding) and outside the image.  Input addressed through strides, so the
// same kernel reads NCHW images and channels-last intermediates.  One thread per 8 outputs.
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const float* __restrict__ in, void* __restrict__ a, int batch, int chans, int hw,
                 long long sb, long long sc, long long sy, long long sx, int kpad, int out_kind,
                 long long plane) {
  const int k8 = kpad >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(batch) * hw * hw * k8;
  if (i >= total) return;
  const int kg = static_cast<int>(i % k8);
  const long long m = i / k8;
  const int x = static_cast<int>(m % hw);
  const int y = static_cast<int>((m / hw) % hw);
  const long long b = m / (static_cast<long long>(hw) * hw);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = kg * 8 + j;
    const int tap = k / chans, c = k - tap * chans;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    v[j] = (tap < 9 && yy >= 0 && yy < hw && xx >= 0 && xx < hw)
               ? __ldg(in + b * sb + c * sc + yy * sy + xx * sx)
               : 0.f;
  }
  const long long o = m * kpad + kg * 8;
  if (out_kind == DEVIT_OUT_BF16) {
    uint4 t;
    t.x = pack_bf16x2(v[0], v[1]);
    t.y = pack_bf16x2(v[2], v[3]);
    t.z = pack_bf16x2(v[4], v[5]);
    t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a) + o) = t;
  } else {
    float* hi = static_cast<float*>(a) + o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float h = out_kind == DEVIT_OUT_F32_SPLIT ? tf32_hi(v[j]) : v[j];
      hi[j] = h;
      if (out_kind == DEVIT_OUT_F32_SPLIT) hi[plane + j] = v[j] - h;
    }
  }
}

// 3x3 / stride 2 / pad 1 max-pool over a channels-last map [B, hw, hw, C] (bf16 or fp32, already
// ReLU'd so every window has a valid non-negative element), fp32 channels-last output
// [B, hw/2, hw/2, C], optionally + pos[(oy*(hw/2) + ox), c].  One thread per 4 channels.
template <bool BF16>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const void* __restrict__ in, float* __restrict__ out,
                    const float* __restrict__ pos, int batch, int hw, int chans) {
  const int c4n = chans >> 2, oh = hw >> 1;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(batch) * oh * oh * c4n;
  if (i >= total) return;
  const int c4 = static_cast<int>(i % c4n);
  long long t = i / c4n;
  const int ox = static_cast<int>(t % oh);
  t /= oh;
  const int oy = static_cast<int>(t % oh);
  const long long b = t / oh;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int py = 0; py < 3; ++py) {
    const int y = 2 * oy - 1 + py;
    if (y < 0 || y >= hw) continue;
#pragma unroll
    for (int px = 0; px < 3; ++px) {
      const int x = 2 * ox - 1 + px;
      if (x < 0 || x >= hw) continue;
      const long long e = ((b * hw + y) * hw + x) * chans + c4 * 4;
      float4 v;
      if (BF16) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(in) + e));
        v.x = __uint_as_float(raw.x << 16);
        v.y = __uint_as_float(raw.x & 0xFFFF0000u);
        v.z = __uint_as_float(raw.y << 16);
        v.w = __uint_as_float(raw.y & 0xFFFF0000u);
      } else {
        v = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(in) + e));
      }
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  if (pos) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + (oy * oh + ox) * c4n + c4);
    m.x += p.x; m.y += p.y; m.z += p.z; m.w += p.w;
  }
  reinterpret_cast<float4*>(out)[i] = m;
}

// Sequence pooling of one image per CTA (256 threads): a_t = xn[t,:].w + b, p = softmax_t(a),
// pooled[d] = sum_t p_t xn[t, d].  xn fp32 [batch, tokens, dim], tokens <= 256, dim % 4 == 0.
__global__ void __launch_bounds__(256)
seqpool_kernel(const float* __restrict__ xn, const float* __restrict__ w, float bias,
               float* __restrict__ pooled, int tokens, int dim) {
  __shared__ float s_a[256];
  __shared__ float s_red[8];
  const float* xi = xn + static_cast<long long>(blockIdx.x) * tokens * dim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < tokens; t += 8) {  // one warp per token: coalesced row reads
    float acc = 0.f;
    for (int d = lane; d < dim; d += 32) acc = fmaf(xi[t * dim + d], __ldg(w + d), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_a[t] = acc + bias;
  }
  __syncthreads();
  float v = threadIdx.x < tokens ? s_a[threadIdx.x] : -INFINITY;
  float mx = v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) mx = fmaxf(mx, s_red[k]);
  __syncthreads();
  const float e = threadIdx.x < tokens ? expf(v - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_red[warp] = sum;
  if (threadIdx.x < tokens) s_a[threadIdx.x] = e;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) sum += s_red[k];
  const float inv = 1.0f / sum;
  for (int d = threadIdx.x; d < dim; d += 256) {
    float acc = 0.f;
    for (int t = 0; t < tokens; ++t) acc = fmaf(s_a[t], xi[t * dim + d], acc);
    pooled[static_cast<long long>(blockIdx.x) * dim + d] = acc * inv;
  }
}

}  // namespace devit

using namespace devit;

extern "C" int devit_im2col3x3(const float* in, void* a, int32_t batch, int32_t chans, int32_t hw,
                               int64_t sb, int64_t sc, int64_t sy, int64_t sx, int32_t kpad,
                               int32_t out_kind, int64_t out_plane_stride, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(in && a, "devit_im2col3x3: null pointer");
  DEVIT_REQUIRE(batch > 0 && chans > 0 && hw > 0 && kpad >= 9 * chans && kpad % 8 == 0,
                "devit_im2col3x3: need kpad >= 9*chans and kpad %% 8 == 0 (got %d for %d chans)",
                kpad, chans);
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(a) % 16 == 0, "devit_im2col3x3: a must be 16-byte aligned");
  const long long total = static_cast<long long>(batch) * hw * hw * (kpad / 8);
  {
    ProfScope ps(kTagIm2col, stream);
    im2col3x3_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        in, a, batch, chans, hw, sb, sc, sy, sx, kpad, out_kind, out_plane_stride);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" int devit_maxpool3x3s2_cl(const void* in, int32_t in_kind, float* out, const float* pos,
                                     int32_t batch, int32_t hw, int32_t chans, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(in && out, "devit_maxpool3x3s2_cl: null pointer");
  DEVIT_REQUIRE(batch > 0 && hw > 0 && hw % 2 == 0 && chans > 0 && chans % 4 == 0,
                "devit_maxpool3x3s2_cl: need an even map side and chans %% 4 == 0");
  DEVIT_REQUIRE(in_kind == DEVIT_OUT_BF16 || in_kind == DEVIT_OUT_F32,
                "devit_maxpool3x3s2_cl: input must be bf16 or fp32");
  const long long total = static_cast<long long>(batch) * (hw / 2) * (hw / 2) * (chans / 4);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  {
    ProfScope ps(kTagPrefix, stream);
    if (in_kind == DEVIT_OUT_BF16)
      maxpool3x3s2_kernel<true><<<grid, 256, 0, stream>>>(in, out, pos, batch, hw, chans);
    else
      maxpool3x3s2_kernel<false><<<grid, 256, 0, stream>>>(in, out, pos, batch, hw, chans);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" int devit_seqpool(const float* xn, const float* w, float b, float* pooled,
                             int32_t batch, int32_t tokens, int32_t dim, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(xn && w && pooled, "devit_seqpool: null pointer");
  DEVIT_REQUIRE(batch > 0 && tokens > 0 && tokens <= 256 && dim > 0,
                "devit_seqpool: need 0 < tokens <= 256 (got %d)", tokens);
  {
    ProfScope ps(kTagGatherLn, stream);
    seqpool_kernel<<<batch, 256, 0, stream>>>(xn, w, b, pooled, tokens, dim);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}
