// Host-side helpers shared by the launchers: error reporting, device check, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/devit_b200.h"

namespace devit {

int set_error(int code, const char* fmt, ...);
int check_device();  // DEVIT_OK iff current device is sm_100
int num_sms();
void count_launch(int n = 1);

// Optional per-launch device timing (devit_profile_enable): when on, each launcher brackets
// its kernel with two CUDA events on the launch stream, tagged by kernel family.
enum ProfTag {
  kTagGemmOther = 0, kTagGemmPatch = 1, kTagGemmQkv = 2, kTagGemmProj = 3, kTagGemmFc1 = 4,
  kTagGemmFc2 = 5, kTagGemmFusion = 6, kTagGemmHead = 7, kTagAttention = 8, kTagLayerNorm = 9,
  kTagGatherLn = 10, kTagIm2col = 11, kTagPrefix = 12, kTagMlpFused = 13, kTagEvalTail = 14,
  kNumProfTags = 16
};
struct ProfScope {
  cudaStream_t stream;
  bool on;
  void* end_event = nullptr;  // this scope's own stop event (launches on other threads / streams
                              // may open scopes in between)
  ProfScope(int tag, cudaStream_t s);
  ~ProfScope();
};

// 1 unless DEVIT_PDL=0: launch the layer-loop kernels with programmatic stream serialization
int pdl_enabled();
// Integer value of a debug / tuning environment variable, read ONCE per process and cached in
// `*cache` (which must start at kEnvUnread); `dflt` when the variable is not set.
constexpr int kEnvUnread = -0x7fffffff;
int env_int(const char* name, int dflt, int* cache);

#define DEVIT_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::devit::set_error(DEVIT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,          \
                                cudaGetErrorString(_e), __FILE__, __LINE__);             \
  } while (0)

#define DEVIT_REQUIRE(cond, ...)                                         \
  do {                                                                   \
    if (!(cond)) return ::devit::set_error(DEVIT_ERR_ARG, __VA_ARGS__);  \
  } while (0)

// 2D / 3D row-major tensor maps with the 128-byte swizzle.  `elem_bytes` is 2 (bf16) or 4
// (fp32, consumed as tf32).  Dimensions are given innermost first; strides in ELEMENTS.
int encode_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t cols,
                   uint64_t rows, uint64_t ld, uint32_t box_cols, uint32_t box_rows,
                   bool weight_like);
// 2D map with the 64-byte swizzle: shared-memory rows of 64 B, 16-byte chunk c of row r stored
// at chunk c ^ ((r >> 1) & 3)
int encode_tmap_2d_sw64(CUtensorMap* map, const void* base, int elem_bytes, uint64_t cols,
                          uint64_t rows, uint64_t ld, uint32_t box_cols, uint32_t box_rows);
int encode_tmap_3d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t d0, uint64_t d1,
                   uint64_t d2, uint64_t stride1, uint64_t stride2, uint32_t box0,
                   uint32_t box1, uint32_t box2);

}  // namespace devit
