// Per-SM global-store throughput on B200: how fast can ONE SM (and all of them together) write
// (a) coalesced 16-byte STG from registers, (b) row-per-lane 32-byte STG.256 (the pattern of a
// TMEM-lane-per-row epilogue), (c) TMA bulk-tensor stores of [32 x 128 B] boxes from shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu -lcuda && ./store_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(512) k_coalesced(uint4* out, long long per_cta_vec, int iters) {
  uint4* base = out + (long long)blockIdx.x * per_cta_vec;
  const uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
  for (int it = 0; it < iters; ++it)
    for (long long i = threadIdx.x; i < per_cta_vec; i += blockDim.x) base[i] = v;
}

// each lane writes its own "row": 32 B pieces at a 1536-byte row stride (fp32 D = 384 rows)
__global__ void __launch_bounds__(512) k_rowlane(uint8_t* out, long long per_cta_bytes, int iters) {
  uint8_t* base = out + (long long)blockIdx.x * per_cta_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = per_cta_bytes / 1536;
  for (int it = 0; it < iters; ++it)
    for (long long r0 = warp * 32; r0 + 32 <= rows; r0 += 16 * 32)
      for (int c = 0; c < 1536; c += 32) {
        void* p = base + (r0 + lane) * 1536 + c;
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(lane) : "memory");
      }
}

__global__ void __launch_bounds__(512) k_tma(const __grid_constant__ CUtensorMap tm, int rows_per_cta, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* slot = smem + warp * 4096;
  for (int i = lane; i < 1024; i += 32) reinterpret_cast<uint32_t*>(slot)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int row_base = blockIdx.x * rows_per_cta;
  for (int it = 0; it < iters; ++it)
    for (int r0 = warp * 32; r0 + 32 <= rows_per_cta; r0 += 16 * 32)
      for (int c = 0; c < 384; c += 32) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"((uint32_t)__cvta_generic_to_shared(slot)),
                         "r"(c), "r"(row_base + r0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        }
        __syncwarp();
      }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const long long per_cta = 3ll << 20;  // 3 MB per CTA per iteration (rows of 1536 B: 2048 rows)
  const int iters = 4;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  uint8_t* buf;
  CK(cudaMalloc(&buf, per_cta * 148));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int rows_per_cta = (int)(per_cta / 1536);
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 4096 + 1024));
  for (int ctas : {1, 16, 74, 148}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {384, (cuuint64_t)rows_per_cta * ctas};
    cuuint64_t strides[1] = {1536};
    cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
      printf("encode failed\n"); return 1;
    }
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) k_coalesced<<<ctas, 512>>>((uint4*)buf, per_cta / 16, iters);
        if (mode == 1) k_rowlane<<<ctas, 512>>>(buf, per_cta, iters);
        if (mode == 2) k_tma<<<ctas, 512, 16 * 4096 + 1024>>>(tm, rows_per_cta, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      CK(cudaGetLastError());
      const double bytes = (double)per_cta * iters * ctas;
      const double gbs = bytes / (best * 1e-3) / 1e9;
      printf("ctas=%3d %-22s %8.3f ms  %8.1f GB/s total  %6.1f GB/s per SM  (~%.1f B/clk/SM at %.2f GHz max clock)\n",
             ctas, mode == 0 ? "coalesced STG.128" : mode == 1 ? "row-per-lane STG.256" : "TMA store 4 KB boxes",
             best, gbs, gbs / ctas, gbs / ctas / (clk_khz * 1e-6), clk_khz * 1e-6);
    }
  }
  return 0;
}
