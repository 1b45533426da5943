// The two steps either side of the forward path, both HBM-bound:
//  * uint8 images -> normalised token-row patch matrix: ToTensor + Normalize
//    (data/get_dataset.py:107-108) + images.to(device) (engine.py:224) folded into the patch
//    im2col, so a batch crosses PCIe as bytes instead of fp32 (4x less) and the fp32 image tensor
//    never exists in HBM;
//  * logits -> (cross-entropy, correct@1, correct@5) accumulated on the device: the eval tail of
//    engine.py:229-238 without a host synchronisation per batch.
#include "common.cuh"
#include "ptx.cuh"

namespace devit {

struct NormParams {
  float mean[4];
  float stdv[4];
};

// (u8 / 255 - mean) / std with IEEE divisions and no contraction: the exact fp32 sequence of
// torchvision's ToTensor (.div(255)) followed by Normalize (.sub_(mean).div_(std)).
__device__ __forceinline__ float norm_px(uint32_t byte, float mean, float stdv) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(byte), 255.0f), mean), stdv);
}

// 32-byte global store (STG.256, sm_100): one full sector per lane and instruction -- two
// 16-byte stores at a 32-byte lane stride would each write half sectors.
__device__ __forceinline__ void st_global_256(void* p, uint32_t a0, uint32_t a1, uint32_t a2,
                                              uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6,
                                              uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}
__device__ __forceinline__ void st_global_256f(float* p, const float* v) {
  st_global_256(p, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                __float_as_uint(v[3]), __float_as_uint(v[4]), __float_as_uint(v[5]),
                __float_as_uint(v[6]), __float_as_uint(v[7]));
}

__device__ __forceinline__ void store16(void* a, long long o, const float (&v)[16], int out_kind,
                                        long long plane) {
  if (out_kind == DEVIT_OUT_BF16) {
    st_global_256(static_cast<__nv_bfloat16*>(a) + o, pack_bf16x2(v[0], v[1]),
                  pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]),
                  pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                  pack_bf16x2(v[14], v[15]));
  } else if (out_kind == DEVIT_OUT_F32) {
    st_global_256f(static_cast<float*>(a) + o, v);
    st_global_256f(static_cast<float*>(a) + o + 8, v + 8);
  } else {
    float h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      h[i] = tf32_hi(v[i]);
      l[i] = v[i] - h[i];
    }
    st_global_256f(static_cast<float*>(a) + o, h);
    st_global_256f(static_cast<float*>(a) + o + 8, h + 8);
    st_global_256f(static_cast<float*>(a) + o + plane, l);
    st_global_256f(static_cast<float*>(a) + o + plane + 8, l + 8);
  }
}

// One item = one 16-pixel patch-row segment (y, x0 .. x0+15).  NCHW: one 16-byte load per
// (channel, segment) item; NHWC (3 channels): three 16-byte loads hold the 48 interleaved bytes of
// the segment and feed three 16-element output runs.  Output runs are contiguous K ranges of a
// patch row (32 B bf16 / 64 B fp32); items are numbered in OUTPUT order (image, patch, channel,
// py), so consecutive threads write consecutive runs of the patch matrix.
// The two IEEE divisions per pixel made a first version compute-bound (30 % of the HBM
// roofline): a pixel can only take 256 values per channel, so every block first builds the
// chans x 256 table of normalised values in shared memory with exactly those operations and then
// only looks bytes up; kU8ItemsPerThread items per thread amortise the table.
constexpr int kU8ItemsPerThread = 4;

template <bool NHWC>
__global__ void __launch_bounds__(256)
im2col16_u8_kernel(const uint8_t* __restrict__ img, void* __restrict__ a, int batch, int chans,
                   int hw, NormParams np, int out_kind, long long plane, int row_off,
                   int rows_per_img) {
  __shared__ float lut[4 * 256];
  for (int c = 0; c < chans; ++c) {
    const float mu = c == 0 ? np.mean[0] : c == 1 ? np.mean[1] : c == 2 ? np.mean[2] : np.mean[3];
    const float sd = c == 0 ? np.stdv[0] : c == 1 ? np.stdv[1] : c == 2 ? np.stdv[2] : np.stdv[3];
    lut[c * 256 + threadIdx.x] = norm_px(threadIdx.x, mu, sd);
  }
  __syncthreads();
  const int g = hw >> 4;
  const int P = g * g;
  const int items_c = NHWC ? 1 : chans;
  const long long total = static_cast<long long>(batch) * P * items_c * 16;
  const long long kdim = static_cast<long long>(chans) * 256;
#pragma unroll 1
  for (int it = 0; it < kU8ItemsPerThread; ++it) {
    const long long i = (static_cast<long long>(blockIdx.x) * kU8ItemsPerThread + it) * 256 +
                        threadIdx.x;
    if (i >= total) return;
    const int py = static_cast<int>(i) & 15;
    long long t = i >> 4;
    const int c0 = NHWC ? 0 : static_cast<int>(t % items_c);
    t /= items_c;
    const int pidx = static_cast<int>(t % P);
    const int b = static_cast<int>(t / P);
    const int gy = pidx / g, gx = pidx - gy * g;
    const int y = gy * 16 + py;
    const long long m = static_cast<long long>(b) * rows_per_img + row_off + pidx;
    const long long o0 = m * kdim + py * 16;
    // 16-byte units: NCHW pixel offset / 16, NHWC pixel offset * 3 / 16
    const long long src16 = NHWC ? ((static_cast<long long>(b) * hw + y) * g + gx) * 3
                                 : ((static_cast<long long>(b) * chans + c0) * hw + y) * g + gx;
    float v[16];
    if (!NHWC) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(img) + src16);
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
      const float* tab = lut + c0 * 256;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = tab[(ws[j >> 2] >> (8 * (j & 3))) & 0xffu];
      store16(a, o0 + static_cast<long long>(c0) * 256, v, out_kind, plane);
    } else {
      // 48 bytes: pixel p, channel c at byte 3 p + c
      const uint4* src = reinterpret_cast<const uint4*>(img) + src16;
      const uint4 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
      const uint32_t ws[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                               w2.x, w2.y, w2.z, w2.w};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int byte = 3 * j + c;
          v[j] = lut[c * 256 + ((ws[byte >> 2] >> (8 * (byte & 3))) & 0xffu)];
        }
        store16(a, o0 + static_cast<long long>(c) * 256, v, out_kind, plane);
      }
    }
  }
}

// ------------------------------------------------------------------------------- eval tail
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per sample: nll[b] = logsumexp(logits[b]) - logits[b, t]  (fp32, max-subtracted like
// torch's log_softmax) and rank[b] = number of classes ordered before the target by a
// descending sort (strictly larger logit, or equal logit with a smaller class index).
__global__ void __launch_bounds__(256)
eval_rows_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ target,
                 int batch, int classes, float* __restrict__ nll, int* __restrict__ rank) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= batch) return;
  const float* row = logits + static_cast<long long>(b) * ld;
  const long long t = target[b];
  const bool valid = t >= 0 && t < classes;
  const float lt = valid ? __ldg(row + t) : 0.f;
  float mx = -INFINITY;
  int ahead = 0;
  for (int j = lane; j < classes; j += 32) {
    const float v = __ldg(row + j);
    mx = fmaxf(mx, v);
    ahead += (v > lt || (v == lt && j < t)) ? 1 : 0;
  }
  mx = warp_max_f(mx);
  ahead = warp_sum_i(ahead);
  float s = 0.f;
  for (int j = lane; j < classes; j += 32) s += expf(__ldg(row + j) - mx);
  s = warp_sum_f(s);
  if (lane == 0) {
    nll[b] = valid ? (logf(s) + mx) - lt : __int_as_float(0x7fc00000);
    rank[b] = valid ? ahead : classes;
  }
}

// One block: fixed-order reduction of the per-sample results (deterministic), then the running
// meters acc[5] += {mean loss of the batch, 1, correct@1, correct@k, batch}  -- the updates
// MetricLogger receives per batch in engine.py:235-238 (loss with n = 1, accuracies with
// n = batch) -- and this batch's {mean loss, correct@1, correct@k} into batch_out.
__global__ void __launch_bounds__(256)
eval_reduce_kernel(const float* __restrict__ nll, const int* __restrict__ rank, int batch, int topk,
                   double* __restrict__ acc, float* __restrict__ batch_out) {
  __shared__ double s_loss[256];
  __shared__ int s_c1[256], s_ck[256];
  double l = 0.0;
  int c1 = 0, ck = 0;
  for (int b = threadIdx.x; b < batch; b += 256) {
    l += static_cast<double>(nll[b]);
    const int r = rank[b];
    c1 += r < 1;
    ck += r < topk;
  }
  s_loss[threadIdx.x] = l;
  s_c1[threadIdx.x] = c1;
  s_ck[threadIdx.x] = ck;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_loss[threadIdx.x] += s_loss[threadIdx.x + o];
      s_c1[threadIdx.x] += s_c1[threadIdx.x + o];
      s_ck[threadIdx.x] += s_ck[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = s_loss[0] / batch;
    if (acc) {
      acc[0] += mean;
      acc[1] += 1.0;
      acc[2] += s_c1[0];
      acc[3] += s_ck[0];
      acc[4] += batch;
    }
    if (batch_out) {
      batch_out[0] = static_cast<float>(mean);
      batch_out[1] = static_cast<float>(s_c1[0]);
      batch_out[2] = static_cast<float>(s_ck[0]);
    }
  }
}

}  // namespace devit

using namespace devit;

extern "C" int devit_im2col_tokens_u8(const uint8_t* images, int32_t layout, const float* mean,
                                      const float* stdv, void* a, int32_t batch, int32_t chans,
                                      int32_t hw, int32_t num_prefix, int32_t out_kind,
                                      int64_t out_plane_stride, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(images && a && mean && stdv, "devit_im2col_tokens_u8: null pointer");
  DEVIT_REQUIRE(layout == DEVIT_LAYOUT_NCHW || layout == DEVIT_LAYOUT_NHWC,
                "devit_im2col_tokens_u8: bad layout %d", layout);
  DEVIT_REQUIRE(batch > 0 && chans > 0 && chans <= 4 && hw > 0 && hw % 16 == 0 && num_prefix >= 0,
                "devit_im2col_tokens_u8: need 1..4 channels and an image side that is a positive "
                "multiple of 16 (got %d channels, side %d)", chans, hw);
  DEVIT_REQUIRE(layout == DEVIT_LAYOUT_NCHW || chans == 3,
                "devit_im2col_tokens_u8: the NHWC layout is implemented for 3 channels");
  DEVIT_REQUIRE(out_kind >= 0 && out_kind <= 2, "devit_im2col_tokens_u8: bad out_kind %d", out_kind);
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(images) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(a) % 32 == 0 &&
                    (out_kind != DEVIT_OUT_F32_SPLIT || (out_plane_stride * 4) % 32 == 0),
                "devit_im2col_tokens_u8: images must be 16-byte, the patch matrix (and its plane "
                "stride) 32-byte aligned");
  NormParams np;
  for (int c = 0; c < 4; ++c) {
    np.mean[c] = c < chans ? mean[c] : 0.f;
    np.stdv[c] = c < chans ? stdv[c] : 1.f;
    DEVIT_REQUIRE(np.stdv[c] != 0.f, "devit_im2col_tokens_u8: std[%d] is zero", c);
  }
  const int tokens = num_prefix + (hw / 16) * (hw / 16);
  const size_t esz = out_kind == DEVIT_OUT_BF16 ? 2 : 4;
  const size_t k = static_cast<size_t>(chans) * 256;
  const bool nhwc = layout == DEVIT_LAYOUT_NHWC;
  const long long total = static_cast<long long>(batch) * (nhwc ? 1 : chans) * hw * (hw / 16);
  const long long per_block = 256LL * kU8ItemsPerThread;
  const unsigned grid = static_cast<unsigned>((total + per_block - 1) / per_block);
  {
    ProfScope ps(kTagIm2col, stream);
    if (num_prefix > 0) {  // zero rows for the cls / dist tokens of every image
      DEVIT_CUDA_OK(cudaMemset2DAsync(a, tokens * k * esz, 0, num_prefix * k * esz, batch, stream));
      if (out_kind == DEVIT_OUT_F32_SPLIT)
        DEVIT_CUDA_OK(cudaMemset2DAsync(static_cast<float*>(a) + out_plane_stride, tokens * k * esz,
                                        0, num_prefix * k * esz, batch, stream));
    }
    if (nhwc)
      im2col16_u8_kernel<true><<<grid, 256, 0, stream>>>(images, a, batch, chans, hw, np, out_kind,
                                                         out_plane_stride, num_prefix, tokens);
    else
      im2col16_u8_kernel<false><<<grid, 256, 0, stream>>>(images, a, batch, chans, hw, np, out_kind,
                                                          out_plane_stride, num_prefix, tokens);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

extern "C" size_t devit_eval_tail_workspace_bytes(int32_t batch) {
  return batch > 0 ? static_cast<size_t>(batch) * 8 : 0;
}

extern "C" int devit_eval_tail(const float* logits, int64_t ld, const int64_t* target,
                               int32_t batch, int32_t classes, int32_t topk, void* workspace,
                               size_t workspace_bytes, double* acc, float* batch_out,
                               void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(logits && target && workspace && (acc || batch_out),
                "devit_eval_tail: null pointer");
  DEVIT_REQUIRE(batch > 0 && classes > 0 && ld >= classes && topk >= 1,
                "devit_eval_tail: bad shape (batch %d, classes %d, topk %d)", batch, classes, topk);
  if (workspace_bytes < devit_eval_tail_workspace_bytes(batch))
    return set_error(DEVIT_ERR_WORKSPACE, "devit_eval_tail: workspace %zu < %zu bytes",
                     workspace_bytes, devit_eval_tail_workspace_bytes(batch));
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 8 == 0 &&
                    (!acc || reinterpret_cast<uintptr_t>(acc) % 8 == 0),
                "devit_eval_tail: workspace / acc must be 8-byte aligned");
  float* nll = static_cast<float*>(workspace);
  int* rank = reinterpret_cast<int*>(nll + batch);
  const int k = topk < classes ? topk : classes;  // timm accuracy(): maxk = min(max(topk), C)
  {
    ProfScope ps(kTagEvalTail, stream);
    eval_rows_kernel<<<(batch + 7) / 8, 256, 0, stream>>>(
        logits, ld, reinterpret_cast<const long long*>(target), batch, classes, nll, rank);
    eval_reduce_kernel<<<1, 256, 0, stream>>>(nll, rank, batch, k, acc, batch_out);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return DEVIT_OK;
}
