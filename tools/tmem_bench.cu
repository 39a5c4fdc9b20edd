// Microbenchmark of TMEM read throughput / latency (tcgen05.ld 32x32b) as seen by the attention kernels:
// W warps of ONE CTA per SM each loop over tcgen05.ld.x16 (+ wait::ld) of their own 32-lane quarter.
//   MODE 0: one x16 load + wait per iteration (latency-exposed, what one compute warp sees)
//   MODE 1: two x16 loads + one wait (the backward's S^T / dP^T pair)
// Reports bytes per clock per SM and clocks per iteration per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tools/tmem_bench.cu && ./tmem_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../object_detection_destr_b200/csrc/sm100_ptx.cuh"

using namespace destr;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* cycles, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) * 64) % 448;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      uint32_t r[16];
      tmem_ld_x16(addr, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= r[i];
    } else if (MODE == 1) {
      uint32_t r[16], s[16];
      tmem_ld_x16(addr, r);
      tmem_ld_x16(addr + 16, s);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= r[i] + s[i];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 0x12345678u) out[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int MODE>
void run(int warps, uint32_t* out, long long* cyc) {
  const int iters = 4096;
  bench<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  bench<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double words = (MODE == 0 ? 16 : 32);
  const double bytes = words * 4 * 32 * warps * iters;
  printf("mode %d  warps %2d: %8.1f clk/iter/warp   %7.1f B/clk/SM   (%s)\n", MODE, warps, double(c) / iters, bytes / c,
         cudaGetErrorString(e));
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 4096);
  cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16, 32}) {
    run<0>(w, out, cyc);
    run<1>(w, out, cyc);
  }
  return 0;
}
