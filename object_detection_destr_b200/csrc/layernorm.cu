// Residual-add + LayerNorm family (bf16 I/O, fp32 statistics, warp-shuffle reductions).
//   y = LN(a + b)                     encoder_block.py:104-110,:40 ; decoder_block.py:65,253-258
//   o = lam*LN1(x+o1) + (1-lam)*LN2(x+o2)   decoder_block.py:182-184
// One warp per row; a lane owns D/32 contiguous channels (one or two 128-bit loads per operand).
// HBM-bound: algorithmic bytes = (2 reads + 1 write) * M * D * 2 B for the forward.
#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

constexpr int kSMs = 148;
constexpr float kEps = 1e-5f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VPL>  // values per lane (8 or 16)
struct Row {
  float v[VPL];
};

template <int VPL>
__device__ __forceinline__ Row<VPL> ld_row(const __nv_bfloat16* base, int lane) {
  Row<VPL> r;
#pragma unroll
  for (int c = 0; c < VPL / 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(base + (c * 32 + lane) * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r.v[c * 8 + 2 * i] = __uint_as_float(w[i] << 16);
      r.v[c * 8 + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  return r;
}
template <int VPL>
__device__ __forceinline__ void st_row(__nv_bfloat16* base, int lane, const Row<VPL>& r) {
#pragma unroll
  for (int c = 0; c < VPL / 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(r.v[c * 8 + 2 * i], r.v[c * 8 + 2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(base + (c * 32 + lane) * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// fp32 per-channel vector (gamma/beta) in the same lane ownership
template <int VPL>
__device__ __forceinline__ Row<VPL> ld_vec(const float* base, int lane) {
  Row<VPL> r;
#pragma unroll
  for (int c = 0; c < VPL / 8; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(base + (c * 32 + lane) * 8);
    const float4 b = *reinterpret_cast<const float4*>(base + (c * 32 + lane) * 8 + 4);
    r.v[c * 8 + 0] = a.x; r.v[c * 8 + 1] = a.y; r.v[c * 8 + 2] = a.z; r.v[c * 8 + 3] = a.w;
    r.v[c * 8 + 4] = b.x; r.v[c * 8 + 5] = b.y; r.v[c * 8 + 6] = b.z; r.v[c * 8 + 7] = b.w;
  }
  return r;
}

// dropout on a row in the lane ownership above: value c*8+i <-> column (c*32+lane)*8+i; one hash per column pair
template <int VPL>
__device__ __forceinline__ void drop_row(Row<VPL>& r, uint32_t seed, uint32_t site, uint32_t row, int lane,
                                         uint32_t thr16, float s) {
#pragma unroll
  for (int c = 0; c < VPL / 8; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t col = (c * 32 + lane) * 8 + 2 * i;
      const uint32_t bits = drop_bits(seed, site, row, col >> 1);
      r.v[c * 8 + 2 * i] = ((bits & 0xFFFFu) >= thr16) ? r.v[c * 8 + 2 * i] * s : 0.f;
      r.v[c * 8 + 2 * i + 1] = ((bits >> 16) >= thr16) ? r.v[c * 8 + 2 * i + 1] * s : 0.f;
    }
}

template <int VPL>
__device__ __forceinline__ void row_stats(const Row<VPL>& x, float& mean, float& rstd) {
  constexpr float invD = 1.0f / (VPL * 32);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += x.v[i];
  mean = warp_sum(s) * invD;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float d = x.v[i] - mean;
    q = fmaf(d, d, q);
  }
  rstd = rsqrtf(warp_sum(q) * invD + kEps);
}

template <int VPL>
__global__ void __launch_bounds__(256)
add_ln_fwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                  const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int lda, int ldb, int ldy,
                  Drop dp) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const Row<VPL> g = ld_vec<VPL>(gamma, lane), be = ld_vec<VPL>(beta, lane);
  const uint32_t seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
  const float ds = drop_scale(dp.thr16);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    Row<VPL> x = ld_row<VPL>(a + (size_t)row * lda, lane);
    if (b) {
      Row<VPL> r2 = ld_row<VPL>(b + (size_t)row * ldb, lane);
      if (dp.thr16) drop_row<VPL>(r2, seed, dp.site, row, lane, dp.thr16, ds);  // y = LN(a + dropout(b))
#pragma unroll
      for (int i = 0; i < VPL; ++i) x.v[i] += r2.v[i];
    }
    float mean, rstd;
    row_stats<VPL>(x, mean, rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) x.v[i] = fmaf((x.v[i] - mean) * rstd, g.v[i], be.v[i]);
    st_row<VPL>(y + (size_t)row * ldy, lane, x);
    if (mean_out && lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

// Two chained LayerNorms in one row pass: y1 = LN1(a + dropout(b)), y2 = LN2(c + y1) -- the tail of an encoder layer,
// norm2(x1 + dropout3(fc2 ..)) followed by the Encoder's shared norm(x + block(x)) (encoder_block.py:108-110, :40).
// y1 enters the second LayerNorm as stored (bf16), so the backward kernels, which recompute xhat from the stored
// tensors, see exactly the forward's values.
template <int VPL>
__global__ void __launch_bounds__(256)
add_ln2_fwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                   const float* __restrict__ gamma1, const float* __restrict__ beta1, __nv_bfloat16* __restrict__ y1,
                   float* __restrict__ mean1, float* __restrict__ rstd1, const __nv_bfloat16* __restrict__ c,
                   const float* __restrict__ gamma2, const float* __restrict__ beta2, __nv_bfloat16* __restrict__ y2,
                   float* __restrict__ mean2, float* __restrict__ rstd2, int M, int lda, int ldb, int ldy1, int ldc,
                   int ldy2, Drop dp) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const Row<VPL> g1 = ld_vec<VPL>(gamma1, lane), b1 = ld_vec<VPL>(beta1, lane);
  const Row<VPL> g2 = ld_vec<VPL>(gamma2, lane), b2 = ld_vec<VPL>(beta2, lane);
  const uint32_t seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
  const float ds = drop_scale(dp.thr16);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    Row<VPL> x = ld_row<VPL>(a + (size_t)row * lda, lane);
    Row<VPL> r2 = ld_row<VPL>(b + (size_t)row * ldb, lane);
    Row<VPL> rc = ld_row<VPL>(c + (size_t)row * ldc, lane);
    if (dp.thr16) drop_row<VPL>(r2, seed, dp.site, row, lane, dp.thr16, ds);
#pragma unroll
    for (int i = 0; i < VPL; ++i) x.v[i] += r2.v[i];
    float mean, rstd;
    row_stats<VPL>(x, mean, rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) x.v[i] = fmaf((x.v[i] - mean) * rstd, g1.v[i], b1.v[i]);
    st_row<VPL>(y1 + (size_t)row * ldy1, lane, x);
    if (lane == 0) {
      mean1[row] = mean;
      rstd1[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) x.v[i] = __bfloat162float(__float2bfloat16(x.v[i])) + rc.v[i];  // y1 as stored
    row_stats<VPL>(x, mean, rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) x.v[i] = fmaf((x.v[i] - mean) * rstd, g2.v[i], b2.v[i]);
    st_row<VPL>(y2 + (size_t)row * ldy2, lane, x);
    if (lane == 0) {
      mean2[row] = mean;
      rstd2[row] = rstd;
    }
  }
}

// dx = rstd * (gy - mean(gy) - xhat * mean(gy*xhat)),  gy = dy*gamma;  dgamma += dy*xhat; dbeta += dy
template <int VPL>
__global__ void __launch_bounds__(256)
add_ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ a,
                  const __nv_bfloat16* __restrict__ b, const float* __restrict__ gamma,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                  __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int M,
                  int lddy, int lda, int ldb, int lddx, float* __restrict__ dbias,
                  const __nv_bfloat16* __restrict__ res_in, __nv_bfloat16* __restrict__ res_out, int ldri, int ldro,
                  Drop dp) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  constexpr int D = VPL * 32;
  constexpr float invD = 1.0f / D;
  __shared__ float red[3][8][D];  // [dgamma|dbeta|dbias][warp][channel]  (<= 48 KB at D=512)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const Row<VPL> g = ld_vec<VPL>(gamma, lane);
  const uint32_t seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
  const float ds = drop_scale(dp.thr16);
  Row<VPL> accg, accb, accx;
#pragma unroll
  for (int i = 0; i < VPL; ++i) accg.v[i] = accb.v[i] = accx.v[i] = 0.f;
  // Two rows per warp iteration: every load of both rows is issued before any arithmetic (a warp has only 3-4 rows in
  // all, so without this the kernel is a chain of exposed memory latencies: 13 us against a 3.5 us bandwidth floor).
  const int stride = gridDim.x * wpb;
  for (int row0 = blockIdx.x * wpb + warp; row0 < M; row0 += 2 * stride) {
    constexpr int NR = 2;
    int rows[NR];
    bool ok[NR];
    Row<VPL> xs[NR], bs[NR], dys[NR], rins[NR];
    float means[NR], rstds[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      rows[k] = row0 + k * stride;
      ok[k] = rows[k] < M;
      const int row = ok[k] ? rows[k] : row0;  // (a valid address; the result is discarded)
      xs[k] = ld_row<VPL>(a + (size_t)row * lda, lane);
      if (b) bs[k] = ld_row<VPL>(b + (size_t)row * ldb, lane);
      dys[k] = ld_row<VPL>(dy + (size_t)row * lddy, lane);
      if (res_out && res_in) rins[k] = ld_row<VPL>(res_in + (size_t)row * ldri, lane);
      means[k] = mean_in[row];
      rstds[k] = rstd_in[row];
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      if (!ok[k]) continue;  // warp-uniform
      const int row = rows[k];
      Row<VPL>& x = xs[k];
      if (b) {
        Row<VPL>& r2 = bs[k];
        if (dp.thr16) drop_row<VPL>(r2, seed, dp.site, row, lane, dp.thr16, ds);
#pragma unroll
        for (int i = 0; i < VPL; ++i) x.v[i] += r2.v[i];
      }
      const Row<VPL>& d = dys[k];
      const float mean = means[k], rstd = rstds[k];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        x.v[i] = (x.v[i] - mean) * rstd;  // xhat
        const float gy = d.v[i] * g.v[i];
        s1 += gy;
        s2 = fmaf(gy, x.v[i], s2);
        accg.v[i] = fmaf(d.v[i], x.v[i], accg.v[i]);
        accb.v[i] += d.v[i];
      }
      s1 = warp_sum(s1) * invD;
      s2 = warp_sum(s2) * invD;
      Row<VPL> o;  // gradient w.r.t. the LayerNorm input (a + dropout(b)): what flows into the residual stream `a`
#pragma unroll
      for (int i = 0; i < VPL; ++i) o.v[i] = rstd * (d.v[i] * g.v[i] - s1 - x.v[i] * s2);
      Row<VPL> ob = o;  // gradient w.r.t. b: the same, through the dropout mask
      if (dp.thr16) drop_row<VPL>(ob, seed, dp.site, row, lane, dp.thr16, ds);
#pragma unroll
      for (int i = 0; i < VPL; ++i) accx.v[i] += ob.v[i];
      st_row<VPL>(dx + (size_t)row * lddx, lane, ob);
      if (res_out) {  // res_out = [res_in +] d(a)  (gradient of the residual stream)
        if (res_in) {
#pragma unroll
          for (int i = 0; i < VPL; ++i) o.v[i] += rins[k].v[i];
        }
        st_row<VPL>(res_out + (size_t)row * ldro, lane, o);
      }
    }
  }
  // block reduction of the parameter gradients, then one atomic per channel per block
#pragma unroll
  for (int c = 0; c < VPL / 8; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[0][warp][(c * 32 + lane) * 8 + i] = accg.v[c * 8 + i];
      red[1][warp][(c * 32 + lane) * 8 + i] = accb.v[c * 8 + i];
      red[2][warp][(c * 32 + lane) * 8 + i] = accx.v[c * 8 + i];
    }
  __syncthreads();
  for (int ch = threadIdx.x; ch < D; ch += blockDim.x) {
    float sg = 0.f, sb = 0.f, sx = 0.f;
    for (int w = 0; w < wpb; ++w) {
      sg += red[0][w][ch];
      sb += red[1][w][ch];
      sx += red[2][w][ch];
    }
    atomicAdd(dgamma + ch, sg);
    atomicAdd(dbeta + ch, sb);
    if (dbias) atomicAdd(dbias + ch, sx);
  }
}


// Backward of add_ln2_fwd in one row pass:  y2 = LN2(c + y1),  y1 = LN1(a + dropout(b))
//   d3  = gradient w.r.t. (c + y1)                     (stored: it is also the gradient of the residual stream c)
//   dxb = gradient w.r.t. b (through b's dropout mask), dsum = the un-masked gradient w.r.t. (a + dropout(b))
//   dgamma2/dbeta2 (outer), dgamma1/dbeta1 (inner), dbias (= column sums of dxb: the bias of the Linear that made b)
// d3 enters the inner LayerNorm as stored (bf16), like the two-kernel path it replaces.
template <int VPL>
__global__ void __launch_bounds__(256)
add_ln2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ c,
                   const __nv_bfloat16* __restrict__ y1, const float* __restrict__ gamma2,
                   const float* __restrict__ mean2, const float* __restrict__ rstd2,
                   const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                   const float* __restrict__ gamma1, const float* __restrict__ mean1, const float* __restrict__ rstd1,
                   __nv_bfloat16* __restrict__ d3_out, __nv_bfloat16* __restrict__ dxb_out,
                   __nv_bfloat16* __restrict__ dsum_out, float* __restrict__ dgamma2, float* __restrict__ dbeta2,
                   float* __restrict__ dgamma1, float* __restrict__ dbeta1, float* __restrict__ dbias, int M, Drop dp) {
  constexpr int D = VPL * 32;
  constexpr float invD = 1.0f / D;
  __shared__ float red[5][8][D];  // [dgamma2|dbeta2|dgamma1|dbeta1|dbias][warp][channel]  (40 KB at D = 256)
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const Row<VPL> g2 = ld_vec<VPL>(gamma2, lane), g1 = ld_vec<VPL>(gamma1, lane);
  const uint32_t seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
  const float ds = drop_scale(dp.thr16);
  Row<VPL> acc[5];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[k].v[i] = 0.f;
  for (int row = blockIdx.x * wpb + warp; row < M; row += gridDim.x * wpb) {
    const size_t ro = (size_t)row * D;
    Row<VPL> xo = ld_row<VPL>(c + ro, lane);
    const Row<VPL> ry1 = ld_row<VPL>(y1 + ro, lane);
    const Row<VPL> d = ld_row<VPL>(dy + ro, lane);
    Row<VPL> xi = ld_row<VPL>(a + ro, lane);
    Row<VPL> rb = ld_row<VPL>(b + ro, lane);
    const float m2 = mean2[row], r2 = rstd2[row], m1 = mean1[row], r1 = rstd1[row];
    // ---- outer LayerNorm ----
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xo.v[i] = (xo.v[i] + ry1.v[i] - m2) * r2;  // xhat
      const float gy = d.v[i] * g2.v[i];
      s1 += gy;
      s2 = fmaf(gy, xo.v[i], s2);
      acc[0].v[i] = fmaf(d.v[i], xo.v[i], acc[0].v[i]);
      acc[1].v[i] += d.v[i];
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    Row<VPL> d3;
#pragma unroll
    for (int i = 0; i < VPL; ++i) d3.v[i] = r2 * (d.v[i] * g2.v[i] - s1 - xo.v[i] * s2);
    st_row<VPL>(d3_out + ro, lane, d3);
#pragma unroll
    for (int i = 0; i < VPL; ++i) d3.v[i] = __bfloat162float(__float2bfloat16(d3.v[i]));  // as stored
    // ---- inner LayerNorm ----
    if (dp.thr16) drop_row<VPL>(rb, seed, dp.site, row, lane, dp.thr16, ds);
    s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xi.v[i] = (xi.v[i] + rb.v[i] - m1) * r1;  // xhat
      const float gy = d3.v[i] * g1.v[i];
      s1 += gy;
      s2 = fmaf(gy, xi.v[i], s2);
      acc[2].v[i] = fmaf(d3.v[i], xi.v[i], acc[2].v[i]);
      acc[3].v[i] += d3.v[i];
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    Row<VPL> o;
#pragma unroll
    for (int i = 0; i < VPL; ++i) o.v[i] = r1 * (d3.v[i] * g1.v[i] - s1 - xi.v[i] * s2);
    if (dsum_out) st_row<VPL>(dsum_out + ro, lane, o);
    if (dp.thr16) drop_row<VPL>(o, seed, dp.site, row, lane, dp.thr16, ds);
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[4].v[i] += o.v[i];
    st_row<VPL>(dxb_out + ro, lane, o);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int cch = 0; cch < VPL / 8; ++cch)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[k][warp][(cch * 32 + lane) * 8 + i] = acc[k].v[cch * 8 + i];
  __syncthreads();
  float* outp[5] = {dgamma2, dbeta2, dgamma1, dbeta1, dbias};
  for (int idx = threadIdx.x; idx < 5 * D; idx += blockDim.x) {
    const int which = idx / D, ch = idx - which * D;
    if (outp[which] == nullptr) continue;
    float sacc = 0.f;
    for (int w2 = 0; w2 < wpb; ++w2) sacc += red[which][w2][ch];
    atomicAdd(outp[which] + ch, sacc);
  }
}

// ---------------------------------------------------------------------------------------------
// out = lam*LN1(x+o1) + (1-lam)*LN2(x+o2eff), D = 512  (decoder_block.py:182-184), where o2eff folds
// the head-group slot masking of PairSelfAttention (pair_self_attention.py:101-105):
//   o2eff[c] = [pairs[row,0]==i]*o2[c] + [pairs[row,1]==i]*o2[512+c],  i = row % Q,  o2 is [M,1024].
// stats[row] = {mean1, rstd1, mean2, rstd2}.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dual_ln_mix_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ o1,
                       const __nv_bfloat16* __restrict__ o2, const int32_t* __restrict__ pairs,
                       const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
                       const float* __restrict__ b2, float lam, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ stats, int M, int Q, Drop dp, uint32_t site2) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  constexpr int VPL = 16, D = 512;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const Row<VPL> G1 = ld_vec<VPL>(g1, lane), B1 = ld_vec<VPL>(b1, lane);
  const Row<VPL> G2 = ld_vec<VPL>(g2, lane), B2 = ld_vec<VPL>(b2, lane);
  const uint32_t seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
  const float ds = drop_scale(dp.thr16);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    const int i = row % Q;
    const float k0 = pairs[2 * row] == i ? 1.f : 0.f, k1 = pairs[2 * row + 1] == i ? 1.f : 0.f;
    const Row<VPL> xr = ld_row<VPL>(x + (size_t)row * D, lane);
    Row<VPL> u = ld_row<VPL>(o1 + (size_t)row * D, lane);
    const Row<VPL> a = ld_row<VPL>(o2 + (size_t)row * 2 * D, lane);
    const Row<VPL> c = ld_row<VPL>(o2 + (size_t)row * 2 * D + D, lane);
    Row<VPL> w;
#pragma unroll
    for (int e = 0; e < VPL; ++e) w.v[e] = k0 * a.v[e] + k1 * c.v[e];
    if (dp.thr16) {  // dropout1(o1), dropout1(o2): two independent masks (decoder_block.py:182-184)
      drop_row<VPL>(u, seed, dp.site, row, lane, dp.thr16, ds);
      drop_row<VPL>(w, seed, site2, row, lane, dp.thr16, ds);
    }
#pragma unroll
    for (int e = 0; e < VPL; ++e) {
      u.v[e] += xr.v[e];
      w.v[e] += xr.v[e];
    }
    float m1, r1, m2, r2;
    row_stats<VPL>(u, m1, r1);
    row_stats<VPL>(w, m2, r2);
    Row<VPL> o;
#pragma unroll
    for (int e = 0; e < VPL; ++e) {
      const float y1 = fmaf((u.v[e] - m1) * r1, G1.v[e], B1.v[e]);
      const float y2 = fmaf((w.v[e] - m2) * r2, G2.v[e], B2.v[e]);
      o.v[e] = lam * y1 + (1.f - lam) * y2;
    }
    st_row<VPL>(out + (size_t)row * D, lane, o);
    if (stats && lane == 0) reinterpret_cast<float4*>(stats)[row] = make_float4(m1, r1, m2, r2);
  }
}

__global__ void __launch_bounds__(256)
dual_ln_mix_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                       const __nv_bfloat16* __restrict__ o1, const __nv_bfloat16* __restrict__ o2,
                       const int32_t* __restrict__ pairs, const float* __restrict__ g1,
                       const float* __restrict__ g2, const float* __restrict__ stats, float lam,
                       __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ do1,
                       __nv_bfloat16* __restrict__ do2, float* __restrict__ dg1, float* __restrict__ db1,
                       float* __restrict__ dg2, float* __restrict__ db2, int M, int Q, int head_major,
                       float* __restrict__ delta1, float* __restrict__ delta2, Drop dp, uint32_t site2) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  constexpr int VPL = 16, D = 512;
  constexpr float invD = 1.0f / D;
  extern __shared__ float red[];  // [4][warps][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const Row<VPL> G1 = ld_vec<VPL>(g1, lane), G2 = ld_vec<VPL>(g2, lane);
  const uint32_t seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
  const float ds = drop_scale(dp.thr16);
  Row<VPL> ag1, ab1, ag2, ab2;
#pragma unroll
  for (int e = 0; e < VPL; ++e) ag1.v[e] = ab1.v[e] = ag2.v[e] = ab2.v[e] = 0.f;
  for (int row = blockIdx.x * wpb + warp; row < M; row += gridDim.x * wpb) {
    const int i = row % Q;
    const float k0 = pairs[2 * row] == i ? 1.f : 0.f, k1 = pairs[2 * row + 1] == i ? 1.f : 0.f;
    const Row<VPL> xr = ld_row<VPL>(x + (size_t)row * D, lane);
    Row<VPL> u = ld_row<VPL>(o1 + (size_t)row * D, lane);
    const Row<VPL> o1v = u;
    const Row<VPL> a = ld_row<VPL>(o2 + (size_t)row * 2 * D, lane);
    const Row<VPL> c = ld_row<VPL>(o2 + (size_t)row * 2 * D + D, lane);
    const Row<VPL> d = ld_row<VPL>(dout + (size_t)row * D, lane);
    const float4 st = reinterpret_cast<const float4*>(stats)[row];
    Row<VPL> w;
#pragma unroll
    for (int e = 0; e < VPL; ++e) w.v[e] = k0 * a.v[e] + k1 * c.v[e];
    if (dp.thr16) {
      drop_row<VPL>(u, seed, dp.site, row, lane, dp.thr16, ds);
      drop_row<VPL>(w, seed, site2, row, lane, dp.thr16, ds);
    }
    float s11 = 0.f, s12 = 0.f, s21 = 0.f, s22 = 0.f;
#pragma unroll
    for (int e = 0; e < VPL; ++e) {
      u.v[e] = (u.v[e] + xr.v[e] - st.x) * st.y;                                   // xhat1
      w.v[e] = (xr.v[e] + w.v[e] - st.z) * st.w;                                   // xhat2
      const float d1 = lam * d.v[e], d2 = (1.f - lam) * d.v[e];
      const float gy1 = d1 * G1.v[e], gy2 = d2 * G2.v[e];
      s11 += gy1; s12 = fmaf(gy1, u.v[e], s12);
      s21 += gy2; s22 = fmaf(gy2, w.v[e], s22);
      ag1.v[e] = fmaf(d1, u.v[e], ag1.v[e]); ab1.v[e] += d1;
      ag2.v[e] = fmaf(d2, w.v[e], ag2.v[e]); ab2.v[e] += d2;
    }
    s11 = warp_sum(s11) * invD; s12 = warp_sum(s12) * invD;
    s21 = warp_sum(s21) * invD; s22 = warp_sum(s22) * invD;
    Row<VPL> r1, r2, rx, ra, rc;
#pragma unroll
    for (int e = 0; e < VPL; ++e) {
      r1.v[e] = st.y * (lam * d.v[e] * G1.v[e] - s11 - u.v[e] * s12);
      r2.v[e] = st.w * ((1.f - lam) * d.v[e] * G2.v[e] - s21 - w.v[e] * s22);
      rx.v[e] = r1.v[e] + r2.v[e];
    }
    if (dp.thr16) {  // gradients w.r.t. o1 / o2 pass through their dropout masks; dx (residual) does not
      drop_row<VPL>(r1, seed, dp.site, row, lane, dp.thr16, ds);
      drop_row<VPL>(r2, seed, site2, row, lane, dp.thr16, ds);
    }
#pragma unroll
    for (int e = 0; e < VPL; ++e) {
      ra.v[e] = k0 * r2.v[e];
      rc.v[e] = k1 * r2.v[e];
    }
    st_row<VPL>(dx + (size_t)row * D, lane, rx);
    if (!head_major) {
      st_row<VPL>(do1 + (size_t)row * D, lane, r1);
      st_row<VPL>(do2 + (size_t)row * 2 * D, lane, ra);
      st_row<VPL>(do2 + (size_t)row * 2 * D + D, lane, rc);
    } else {
      // head-major gradients for the attention backward: do1 [B,8,Q,64], do2 [B,8,Q,128]
      const int b = row / Q;
      auto st8v = [](__nv_bfloat16* p, const float* v) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          w[e] = *reinterpret_cast<uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
      };
#pragma unroll
      for (int g = 0; g < 2; ++g) {  // lane owns channels (g*32 + lane)*8 .. +8
        const int h1 = g * 4 + (lane >> 3), h2 = g * 2 + (lane >> 4);
        st8v(do1 + ((size_t)(b * 8 + h1) * Q + i) * 64 + (lane & 7) * 8, &r1.v[g * 8]);
        st8v(do2 + ((size_t)(b * 8 + h2) * Q + i) * 128 + (lane & 15) * 8, &ra.v[g * 8]);
        st8v(do2 + ((size_t)(b * 8 + 4 + h2) * Q + i) * 128 + (lane & 15) * 8, &rc.v[g * 8]);
        if (delta1) {  // delta = rowsum(dO o O) per head, for the attention backward
          float p1 = 0.f, pa = 0.f, pc = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            p1 = fmaf(r1.v[g * 8 + e], o1v.v[g * 8 + e], p1);
            pa = fmaf(ra.v[g * 8 + e], a.v[g * 8 + e], pa);
            pc = fmaf(rc.v[g * 8 + e], c.v[g * 8 + e], pc);
          }
#pragma unroll
          for (int off = 1; off < 8; off <<= 1) {
            p1 += __shfl_xor_sync(0xffffffffu, p1, off);
            pa += __shfl_xor_sync(0xffffffffu, pa, off);
            pc += __shfl_xor_sync(0xffffffffu, pc, off);
          }
          pa += __shfl_xor_sync(0xffffffffu, pa, 8);
          pc += __shfl_xor_sync(0xffffffffu, pc, 8);
          if ((lane & 7) == 0) delta1[(size_t)(b * 8 + h1) * Q + i] = p1;
          if ((lane & 15) == 0) {
            delta2[(size_t)(b * 8 + h2) * Q + i] = pa;
            delta2[(size_t)(b * 8 + 4 + h2) * Q + i] = pc;
          }
        }
      }
    }
  }
  float* R[4] = {red, red + wpb * D, red + 2 * wpb * D, red + 3 * wpb * D};
#pragma unroll
  for (int c = 0; c < VPL / 8; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = (c * 32 + lane) * 8 + e;
      R[0][warp * D + ch] = ag1.v[c * 8 + e];
      R[1][warp * D + ch] = ab1.v[c * 8 + e];
      R[2][warp * D + ch] = ag2.v[c * 8 + e];
      R[3][warp * D + ch] = ab2.v[c * 8 + e];
    }
  __syncthreads();
  float* outp[4] = {dg1, db1, dg2, db2};
  for (int idx = threadIdx.x; idx < 4 * D; idx += blockDim.x) {
    const int which = idx / D, ch = idx - which * D;
    float s = 0.f;
    for (int w2 = 0; w2 < wpb; ++w2) s += R[which][w2 * D + ch];
    atomicAdd(outp[which] + ch, s);
  }
}

inline int ln_grid(int M) {
  int blocks = ceil_div(M, 8);
  const int cap = kSMs * 4;
  return blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace destr

using namespace destr;

#define DESTR_DROP(seed, thr, site) destr::Drop{(seed), (thr), (site)}

extern "C" int destr_add_layernorm_fwd(const void* a, int lda, const void* b, int ldb, const float* gamma,
                                       const float* beta, void* y, int ldy, float* mean, float* rstd, int M, int D,
                                       const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site,
                                       void* stream) {
  DESTR_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldy % 8 == 0 && lda >= D && ldy >= D, "row pitch");
  DESTR_CHECK_ARG(a && gamma && beta && y && M > 0, "null pointer / shape");
  DESTR_CHECK_ARG(D == 256 || D == 512, "D must be 256 or 512");
  DESTR_CHECK_ARG((mean == nullptr) == (rstd == nullptr), "mean and rstd go together");
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 256)
    DESTR_CUDA(launch_k(add_ln_fwd_kernel<8>, dim3(ln_grid(M)), dim3(256), 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, gamma, beta,
                                                     (__nv_bfloat16*)y, mean, rstd, M, lda, ldb, ldy,
                                                     DESTR_DROP(drop_seed, drop_thr16, drop_site)));
  else
    DESTR_CUDA(launch_k(add_ln_fwd_kernel<16>, dim3(ln_grid(M)), dim3(256), 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, gamma, beta,
                                                      (__nv_bfloat16*)y, mean, rstd, M, lda, ldb, ldy,
                                                      DESTR_DROP(drop_seed, drop_thr16, drop_site)));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_add_layernorm2_fwd(const void* a, int lda, const void* b, int ldb, const float* gamma1,
                                        const float* beta1, void* y1, int ldy1, float* mean1, float* rstd1,
                                        const void* c, int ldc, const float* gamma2, const float* beta2, void* y2,
                                        int ldy2, float* mean2, float* rstd2, int M, int D, const uint32_t* drop_seed,
                                        uint32_t drop_thr16, uint32_t drop_site, void* stream) {
  DESTR_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && ldy1 % 8 == 0 && ldy2 % 8 == 0, "row pitch");
  DESTR_CHECK_ARG(a && b && c && gamma1 && beta1 && gamma2 && beta2 && y1 && y2 && mean1 && rstd1 && mean2 && rstd2 && M > 0,
                  "null pointer / shape");
  DESTR_CHECK_ARG(D == 256, "D must be 256");
  DESTR_CUDA(launch_k(add_ln2_fwd_kernel<8>, dim3(ln_grid(M)), dim3(256), 0, (cudaStream_t)stream,
                      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, gamma1, beta1, (__nv_bfloat16*)y1, mean1, rstd1,
                      (const __nv_bfloat16*)c, gamma2, beta2, (__nv_bfloat16*)y2, mean2, rstd2, M, lda, ldb, ldy1, ldc, ldy2,
                      DESTR_DROP(drop_seed, drop_thr16, drop_site)));
  return 0;
}

extern "C" int destr_add_layernorm2_bwd(const void* dy, const void* c, const void* y1, const float* gamma2,
                                        const float* mean2, const float* rstd2, const void* a, const void* b,
                                        const float* gamma1, const float* mean1, const float* rstd1, void* d3,
                                        void* dxb, void* dsum, float* dgamma2, float* dbeta2, float* dgamma1,
                                        float* dbeta1, float* dbias, int M, int D, const uint32_t* drop_seed,
                                        uint32_t drop_thr16, uint32_t drop_site, void* stream) {
  DESTR_CHECK_ARG(dy && c && y1 && gamma2 && mean2 && rstd2 && a && b && gamma1 && mean1 && rstd1 && d3 && dxb && M > 0,
                  "null pointer / shape");
  DESTR_CHECK_ARG(dgamma2 && dbeta2 && dgamma1 && dbeta1, "null parameter-gradient pointer");
  DESTR_CHECK_ARG(D == 256, "D must be 256 (dense [M,256] operands)");
  int grid = ln_grid(M);
  if (grid > kSMs * 2) grid = kSMs * 2;
  DESTR_CUDA(launch_k(add_ln2_bwd_kernel<8>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)dy,
                      (const __nv_bfloat16*)c, (const __nv_bfloat16*)y1, gamma2, mean2, rstd2, (const __nv_bfloat16*)a,
                      (const __nv_bfloat16*)b, gamma1, mean1, rstd1, (__nv_bfloat16*)d3, (__nv_bfloat16*)dxb,
                      (__nv_bfloat16*)dsum, dgamma2, dbeta2, dgamma1, dbeta1, dbias, M,
                      DESTR_DROP(drop_seed, drop_thr16, drop_site)));
  return 0;
}

extern "C" int destr_add_layernorm_bwd(const void* dy, int lddy, const void* a, int lda, const void* b, int ldb,
                                       const float* gamma, const float* mean, const float* rstd, void* dx, int lddx,
                                       float* dgamma, float* dbeta, float* dbias, const void* res_in, int ldri,
                                       void* res_out, int ldro, int M, int D, const uint32_t* drop_seed,
                                       uint32_t drop_thr16, uint32_t drop_site, void* stream) {
  DESTR_CHECK_ARG(lddy % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lddx % 8 == 0 && ldri % 8 == 0 && ldro % 8 == 0,
                  "row pitch");
  DESTR_CHECK_ARG(res_in == nullptr || res_out != nullptr, "res_in needs res_out");
  DESTR_CHECK_ARG(dy && a && gamma && mean && rstd && dx && dgamma && dbeta && M > 0, "null pointer / shape");
  DESTR_CHECK_ARG(D == 256 || D == 512, "D must be 256 or 512");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(M) > kSMs * 2 ? kSMs * 2 : ln_grid(M);
  const __nv_bfloat16* ri = (const __nv_bfloat16*)res_in;
  __nv_bfloat16* ro = (__nv_bfloat16*)res_out;
  if (D == 256)
    DESTR_CUDA(launch_k(add_ln_bwd_kernel<8>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)a,
                                               (const __nv_bfloat16*)b, gamma, mean, rstd, (__nv_bfloat16*)dx, dgamma,
                                               dbeta, M, lddy, lda, ldb, lddx, dbias, ri, ro, ldri, ldro,
                                               DESTR_DROP(drop_seed, drop_thr16, drop_site)));
  else
    DESTR_CUDA(launch_k(add_ln_bwd_kernel<16>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)a,
                                                (const __nv_bfloat16*)b, gamma, mean, rstd, (__nv_bfloat16*)dx, dgamma,
                                                dbeta, M, lddy, lda, ldb, lddx, dbias, ri, ro, ldri, ldro,
                                                DESTR_DROP(drop_seed, drop_thr16, drop_site)));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_dual_ln_mix_fwd(const void* x, const void* o1, const void* o2, const int32_t* pairs,
                                     const float* g1, const float* b1, const float* g2, const float* b2, float lam,
                                     void* out, float* stats, int M, int Q, const uint32_t* drop_seed,
                                     uint32_t drop_thr16, uint32_t drop_site1, uint32_t drop_site2, void* stream) {
  DESTR_CHECK_ARG(x && o1 && o2 && pairs && g1 && b1 && g2 && b2 && out && M > 0 && Q > 0, "null pointer / shape");
  DESTR_CUDA(launch_k(dual_ln_mix_fwd_kernel, dim3(ln_grid(M)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)o1, (const __nv_bfloat16*)o2, pairs, g1, b1, g2, b2, lam,
      (__nv_bfloat16*)out, stats, M, Q, DESTR_DROP(drop_seed, drop_thr16, drop_site1), drop_site2));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_dual_ln_mix_bwd(const void* dout, const void* x, const void* o1, const void* o2,
                                     const int32_t* pairs, const float* g1, const float* g2, const float* stats,
                                     float lam, void* dx, void* do1, void* do2, float* dg1, float* db1, float* dg2,
                                     float* db2, int M, int Q, int head_major, float* delta1, float* delta2,
                                     const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site1,
                                     uint32_t drop_site2, void* stream) {
  DESTR_CHECK_ARG(dout && x && o1 && o2 && pairs && g1 && g2 && stats && dx && do1 && do2 && dg1 && db1 && dg2 && db2,
                  "null pointer");
  DESTR_CHECK_ARG((delta1 == nullptr) == (delta2 == nullptr) && (!delta1 || head_major), "delta needs head_major");
  const int threads = 128;  // 4 warps -> 4*4*512*4 B = 32 KB of reduction scratch
  int grid = ceil_div(M, 4);
  if (grid > kSMs * 2) grid = kSMs * 2;
  DESTR_CUDA(launch_k(dual_ln_mix_bwd_kernel, dim3(grid), dim3(threads), 4 * 4 * 512 * sizeof(float), (cudaStream_t)stream, 
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)x, (const __nv_bfloat16*)o1, (const __nv_bfloat16*)o2, pairs,
      g1, g2, stats, lam, (__nv_bfloat16*)dx, (__nv_bfloat16*)do1, (__nv_bfloat16*)do2, dg1, db1, dg2, db2, M, Q,
      head_major, delta1, delta2, DESTR_DROP(drop_seed, drop_thr16, drop_site1), drop_site2));
  DESTR_LAUNCH_CHECK();
  return 0;
}
