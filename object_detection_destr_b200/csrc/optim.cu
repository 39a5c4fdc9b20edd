// Flat AdamW over the hot path's parameter buffer: ONE launch updates every encoder/decoder parameter
// (fp32 master, exp_avg, exp_avg_sq) and refreshes the bf16 weight shadows the GEMMs read, instead of
// torch's multi-tensor AdamW (8 launches over 102 tensors) plus a separate master -> shadow cast.
// Arithmetic = torch.optim.AdamW (decoupled weight decay, bias correction; the reference trains with
// torch.optim.AdamW, src/train/train.py:236-244):
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The weight-matrix gradients of parameters [bf16_begin, n_bf16) are bf16 (library dW GEMMs, or a bf16 gradient exchange;
// the tcgen05 split-K dW kernel accumulates fp32 straight into `grad`): the kernel reads them
// directly (times grad_scale = 1/world for the data-parallel mean) and writes the fp32 value back to `grad`, which
// replaces a separate bf16 -> fp32 cast pass over the whole buffer (61 us of a 5 ms step).
// HBM-bound: 14-16 B read + 14-18 B written per parameter; 128-bit accesses, grid = multiple of the SM count.
#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

__global__ void __launch_bounds__(256)
flat_adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  __nv_bfloat16* __restrict__ shadow, int64_t n4, float lr, float b1, float b2, float eps, float wd,
                  const float* __restrict__ step, const __nv_bfloat16* __restrict__ g16, int64_t lo16_4, int64_t n16_4,
                  float gscale) {
  const float t = *step;
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg;
    if (i >= lo16_4 && i < n16_4) {  // bf16 gradient of a weight matrix: widen, scale, and leave the fp32 value in .grad
      const uint2 w = reinterpret_cast<const uint2*>(g16)[i];
      gg = make_float4(__uint_as_float(w.x << 16) * gscale, __uint_as_float(w.x & 0xffff0000u) * gscale,
                       __uint_as_float(w.y << 16) * gscale, __uint_as_float(w.y & 0xffff0000u) * gscale);
      reinterpret_cast<float4*>(g)[i] = gg;
    } else {
      gg = reinterpret_cast<const float4*>(g)[i];
      if (gscale != 1.f) {
        gg = make_float4(gg.x * gscale, gg.y * gscale, gg.z * gscale, gg.w * gscale);
        reinterpret_cast<float4*>(g)[i] = gg;
      }
    }
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k];
      const float pk = pa[k] * decay;
      ma[k] = b1 * ma[k] + (1.f - b1) * gk;
      va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
      pa[k] = pk - step_size * (ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps));
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    __nv_bfloat162 lo = __floats2bfloat162_rn(pa[0], pa[1]), hi = __floats2bfloat162_rn(pa[2], pa[3]);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(shadow)[i] = o;
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_flat_adamw(float* master, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                                int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                const float* step, const void* grad_bf16, int64_t bf16_begin, int64_t n_bf16,
                                float grad_scale, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(master && grad && exp_avg && exp_avg_sq && shadow_bf16 && step, "null pointer");
  DESTR_CHECK_ARG(n > 0 && n % 4 == 0, "n must be a positive multiple of 4");
  DESTR_CHECK_ARG(n_bf16 >= 0 && n_bf16 <= n && n_bf16 % 4 == 0 && (n_bf16 == 0 || grad_bf16), "n_bf16 / grad_bf16");
  DESTR_CHECK_ARG(bf16_begin >= 0 && bf16_begin % 4 == 0 && bf16_begin <= n_bf16, "bf16_begin");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  flat_adamw_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      master, grad, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(shadow_bf16), n4, lr, beta1, beta2, eps,
      weight_decay, step, static_cast<const __nv_bfloat16*>(grad_bf16), bf16_begin / 4, n_bf16 / 4, grad_scale);
  DESTR_LAUNCH_CHECK();
  return 0;
}
