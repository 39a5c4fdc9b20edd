// Host-side helpers shared by every translation unit of libdestr_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <utility>

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

int make_tmap_f32_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols, CUtensorMapSwizzle swizzle);

// opt-in to `bytes` of dynamic shared memory for `kernel` on the current device (cached per kernel AND device)
int ensure_dyn_smem(const void* kernel, size_t bytes);
#define DESTR_SMEM_OPTIN(kernel, bytes)                                             \
  do {                                                                              \
    int _rc = ::destr::ensure_dyn_smem(reinterpret_cast<const void*>(kernel), (bytes)); \
    if (_rc) return _rc;                                                            \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL).  A training step is a chain of several hundred SMALL dependent kernels whose
// lifetimes are dominated by fixed costs (grid launch, barrier init, TMEM allocation, descriptor prefetch, pipeline
// fill).  Kernels launched through launch_k() carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may
// be scheduled while the previous kernel of the stream is still draining, run their prologue, and block in
// pdl_wait() (griddepcontrol.wait) until that kernel has completed and its memory is visible; every such kernel calls
// pdl_launch() (griddepcontrol.launch_dependents) once its own prologue is done so that its successor can do the same.
// Rule for kernels: NO global memory access before pdl_wait().  Under stream capture the dependency becomes a
// programmatic edge of the CUDA graph.  DESTR_PDL=0 in the environment (or destr_debug_knob(16, 0)) disables the
// attribute; the device-side instructions are then no-ops.
bool pdl_enabled();  // api.cu
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace destr

// ------------------------------------------------------------------------------------------------------
// Dropout masks: counter-based, no state.  keep(seed, site, row, col) comes from one 32-bit mix hash (of a pre-mixed (seed, site) base, the row and the column pair) per PAIR of
// adjacent columns (even col -> low 16 bits, odd col -> high 16 bits); an element is KEPT when its 16 bits are
// >= thr16 = round(p * 65536), and kept values are scaled by 65536 / (65536 - thr16).  `site` identifies the
// dropout call (layer, position in the layer), `row`/`col` the element; forward and backward kernels recompute the
// same mask from the same (seed, site), whatever their tiling.  oracle/dropout_mask.py is the numpy twin.
// ------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace destr {
// see launch_k(): wait for the previous kernel of the stream (no-op for a normal launch) / let the next one start launching
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// (seed, site) -> avalanche-mixed base: consecutive seeds (the per-step counter) and neighbouring sites give unrelated
// tables, not XOR re-indexings of one another.  Loop-invariant in every kernel (hoisted by the compiler).
__device__ __forceinline__ uint32_t drop_base(uint32_t seed, uint32_t site) {
  uint32_t h = seed * 0x9E3779B1u + site * 0x7F4A7C15u + 0x165667B1u;
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t drop_bits(uint32_t seed, uint32_t site, uint32_t row, uint32_t col_pair) {
  uint32_t h = drop_base(seed, site);
  h ^= row * 0x85EBCA77u;
  h ^= col_pair * 0xC2B2AE3Du;
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
// keep flag of element (row, col)
__device__ __forceinline__ bool drop_keep(uint32_t seed, uint32_t site, uint32_t row, uint32_t col, uint32_t thr16) {
  const uint32_t b = drop_bits(seed, site, row, col >> 1);
  return ((col & 1u) ? (b >> 16) : (b & 0xFFFFu)) >= thr16;
}
__host__ __device__ __forceinline__ float drop_scale(uint32_t thr16) { return 65536.f / (65536.f - (float)thr16); }
// every dropout-capable entry point takes (const uint32_t* seed, uint32_t thr16, uint32_t site): `seed` is a DEVICE
// pointer (so a CUDA graph replays with a fresh seed each step), thr16 = 0 disables dropout
struct Drop {
  const uint32_t* seed;
  uint32_t thr16, site;
};
// the encoder attention kernels take the mask as a precomputed bit matrix instead (attn_dropout_bits.cu)
struct DropBits {
  const uint32_t* bits;  // [B*heads][words][Np], 1 = dropped
  uint32_t thr16;
  int words;
};
}  // namespace destr
#endif
