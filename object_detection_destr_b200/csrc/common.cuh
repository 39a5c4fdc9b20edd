// Host-side helpers shared by every translation unit of libdestr_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace destr {

void set_error(const std::string& msg);  // api.cu

#define DESTR_CHECK_ARG(cond, msg)                                                  \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::destr::set_error(std::string(__func__) + ": bad argument: " + (msg));       \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

#define DESTR_CUDA(expr)                                                            \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::destr::set_error(std::string(__func__) + ": " #expr ": " + cudaGetErrorString(_e)); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

#define DESTR_LAUNCH_CHECK() DESTR_CUDA(cudaGetLastError())

// 2-D bf16 tensor map: `rows` x `cols` elements, row pitch `ld` elements, box = box_rows x box_cols,
// swizzle = CU_TENSOR_MAP_SWIZZLE_{64B,128B}; out-of-bounds elements read as zero.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace destr
