// libdestr_b200.so: version / error plumbing and the TMA tensor-map helper.
#include "../../include/destr_b200.h"
#include "common.cuh"

#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

namespace destr {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  // resolved through the runtime so the library does not link libcuda.so directly
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available (driver too old or no GPU)");
    return 1;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15)) {
    set_error("TMA operand must be 16-byte aligned with a row pitch that is a multiple of 8 elements");
    return 2;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

// 2-D fp32 tensor map (used as the destination of TMA reduce-adds)
int make_tmap_f32_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available (driver too old or no GPU)");
    return 1;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 4) & 15)) {
    set_error("TMA operand must be 16-byte aligned with a row pitch that is a multiple of 4 floats");
    return 2;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute belongs to (kernel, device): a process-wide
// "done" flag would leave a second GPU of the same process without it, so the cache is keyed by both.
int ensure_dyn_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return 1;
  }
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = granted[{kernel, dev}];
  if (bytes <= 48 * 1024 || bytes <= have) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): ") + cudaGetErrorString(e));
    return 1;
  }
  have = bytes;
  return 0;
}

extern int g_knobs[24];  // enc_attn_fwd.cu; knob 16: -1 = read DESTR_PDL from the environment, 0 = off, 1 = on
bool pdl_enabled() {
  static int env = -1;
  if (g_knobs[16] == 0 || g_knobs[16] == 1) return g_knobs[16] == 1;
  if (env < 0) {
    const char* e = getenv("DESTR_PDL");
    env = (e && e[0] == '1') ? 1 : 0;  // default OFF: measured 4.548 ms/step with it vs 4.543 without (the graph's
                                      // kernel-to-kernel hand-over is already cheaper than the overlapped prologues)
  }
  return env == 1;
}

}  // namespace destr

extern "C" int destr_version(void) { return 100; }
extern "C" const char* destr_last_error(void) { return destr::g_last_error.c_str(); }
