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
int encode_tmap_3d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t d0, uint64_t d1,
                   uint64_t d2, uint64_t stride1, uint64_t stride2, uint32_t box0,
                   uint32_t box1, uint32_t box2);

}  // namespace devit
