// Microbenchmark of the exponential throughput that bounds the d_head = 32 encoder attention:
// per-SM rates of MUFU ex2 (f32), packed ex2 (bf16x2), the FMA-pipe polynomial exp2 used by the attention kernels
// (f32x2 Cody-Waite + cubic), and plain / packed FMA for reference.  Full occupancy (2048 threads per SM), 8 independent
// chains per thread, cycles from clock64 -> results per clock per SM, independent of the SM frequency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_bench tools/exp_bench.cu && ./exp_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../object_detection_destr_b200/csrc/sm100_ptx.cuh"

using namespace destr;

__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, long long* cycles, int iters) {
  float a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = -0.001f * (threadIdx.x + i + 1);
    u[i] = 0xBF80BF00u + i;  // two negative bf16 values
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2_approx(a[i]) - 1.0f;                 // 1 MUFU + 1 FADD (keeps the value in range)
      if (MODE == 1) u[i] = ex2_bf16x2(u[i]) ^ 0x80008000u;          // 1 MUFU (2 results) + 1 LOP
      if (MODE == 3) a[i] = fmaf(a[i], 0.999f, 0.001f);
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float p0, p1;
        ex2_poly_f32x2(a[i], a[i + 1], p0, p1);
        a[i] = p0 - 1.0f;
        a[i + 1] = p1 - 1.0f;
      }
    }
    if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        uint64_t v = fma_f32x2(pack_f32x2(a[i], a[i + 1]), pack_f32x2(0.999f, 0.999f), pack_f32x2(0.001f, 0.001f));
        unpack_f32x2(v, a[i], a[i + 1]);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int results_per_chain_step) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, iters = 4096;
  float* out;
  long long* cyc;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaMalloc(&cyc, blocks * sizeof(long long));
  bench<MODE><<<blocks, 256>>>(out, cyc, iters);
  bench<MODE><<<blocks, 256>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long* h = new long long[blocks];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  // per SM: 8 blocks x 256 threads x 8 chains x iters steps, each step = results_per_chain_step results
  const double results = 8.0 * 256 * 8 * iters * results_per_chain_step;
  printf("%-44s %8.1f results / clk / SM   (%.0f cycles for %d iterations)\n", name, results / avg, avg, iters);
  cudaFree(out);
  cudaFree(cyc);
  delete[] h;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("ex2.approx.ftz.f32 (MUFU) + FADD", 1);
  run<1>("ex2.approx.ftz.bf16x2 (MUFU, packed) + LOP", 2);
  run<2>("polynomial exp2 on the FMA pipe (f32x2)", 1);
  run<3>("fma.rn.f32", 1);
  run<4>("fma.rn.f32x2 (packed)", 1);
  return 0;
}
