// fp32 box kernels: pairing (_get_pairs), box refinement, and the block-diagonal matching cost.
// Index / assignment parity with the reference requires its exact fp32 operation order, so every
// product/sum below uses __f*_rn intrinsics (never contracted into FMA by the compiler).
#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

struct XYXY {
  float x0, y0, x1, y1;
};

// from_cxcyhw_to_xyxy (bbox_utils.py:50-63): order (cx, cy, h, w); mins clipped >= 0, maxs <= 1
__device__ __forceinline__ XYXY to_xyxy(float cx, float cy, float hh, float ww) {
  XYXY r;
  r.x0 = fmaxf(__fsub_rn(cx, __fdiv_rn(ww, 2.f)), 0.f);
  r.y0 = fmaxf(__fsub_rn(cy, __fdiv_rn(hh, 2.f)), 0.f);
  r.x1 = fminf(__fadd_rn(cx, __fdiv_rn(ww, 2.f)), 1.f);
  r.y1 = fminf(__fadd_rn(cy, __fdiv_rn(hh, 2.f)), 1.f);
  return r;
}
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// ---------------------------------------------------------------------------------------------
// _get_pairs (pair_self_attention.py:110-171): one WARP per box i, lanes over the candidates j, then a warp
// arg-max with torch.argmax's rule (first maximal element; NaN counts as maximal).  grid = (ceil(Q/8), B).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pair_indices_kernel(const float* __restrict__ coords, int32_t* __restrict__ pairs, int Q) {
  extern __shared__ float sh[];  // [Q][4] xyxy, [Q] area, [Q] l1
  float* bx = sh;
  float* area = sh + 4 * Q;
  float* l1 = area + Q;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < Q; i += blockDim.x) {
    const float4 c = reinterpret_cast<const float4*>(coords)[(size_t)b * Q + i];
    const XYXY r = to_xyxy(c.x, c.y, c.z, c.w);
    bx[4 * i + 0] = r.x0; bx[4 * i + 1] = r.y0; bx[4 * i + 2] = r.x1; bx[4 * i + 3] = r.y1;
    const float w = __fsub_rn(r.x1, r.x0), h = __fsub_rn(r.y1, r.y0);
    area[i] = __fmul_rn(w, h);
    l1[i] = __fadd_rn(fabsf(w), fabsf(h));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= Q) return;  // warp-uniform
  const float x0 = bx[4 * i], y0 = bx[4 * i + 1], x1 = bx[4 * i + 2], y1 = bx[4 * i + 3], ai = area[i];
  // better(a, b): a precedes b in torch.argmax's order (NaN first, then larger value, then smaller index)
  auto better = [](float va, int ja, float vb, int jb) {
    if (jb < 0) return ja >= 0;
    if (ja < 0) return false;
    const bool na = isnan(va), nb = isnan(vb);
    if (na != nb) return na;
    if (!na && va != vb) return va > vb;
    return ja < jb;
  };
  float best = 0.f;
  int bj = -1;
  for (int j = lane; j < Q; j += 32) {
    // unclamped intersection (pair_self_attention.py:122-126)
    const float iw = __fsub_rn(fminf(x1, bx[4 * j + 2]), fmaxf(x0, bx[4 * j + 0]));
    const float ih = __fsub_rn(fminf(y1, bx[4 * j + 3]), fmaxf(y0, bx[4 * j + 1]));
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(ai, area[j]), inter);
    float v = __fdiv_rn(inter, __fadd_rn(uni, 1e-6f));
    v = __fsub_rn(v, (i == j) ? 1.f : 0.f);
    if (better(v, j, best, bj)) { best = v; bj = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (better(ov, oj, best, bj)) { best = ov; bj = oj; }
  }
  if (lane == 0) {
    const bool keep = l1[i] >= l1[bj];
    pairs[((size_t)b * Q + i) * 2 + 0] = keep ? i : bj;
    pairs[((size_t)b * Q + i) * 2 + 1] = keep ? bj : i;
  }
}

// sigmoid(delta + [logit(cx), logit(cy), 0, 0]); logit(x) = -log(1/max(x,1e-6) - 1)  (misc.py:59-62)
__global__ void box_refine_kernel(const float* __restrict__ delta, const float* __restrict__ centers,
                                  float* __restrict__ boxes, int M) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const float4 d = reinterpret_cast<const float4*>(delta)[r];
  const float2 c = reinterpret_cast<const float2*>(centers)[r];
  const float lx = -logf(__fsub_rn(__fdiv_rn(1.f, fmaxf(c.x, 1e-6f)), 1.f));
  const float ly = -logf(__fsub_rn(__fdiv_rn(1.f, fmaxf(c.y, 1e-6f)), 1.f));
  float4 o;
  o.x = 1.f / (1.f + expf(-(d.x + lx)));
  o.y = 1.f / (1.f + expf(-(d.y + ly)));
  o.z = 1.f / (1.f + expf(-d.z));
  o.w = 1.f / (1.f + expf(-d.w));
  reinterpret_cast<float4*>(boxes)[r] = o;
}

// The output layer of the box head (Linear 256 -> 4, fp32 weights: decoder_block.py:51, model.py:33-39) fused with the
// refinement above: boxes = sigmoid(hidden W2^T + b2 + [logit(cx), logit(cy), 0, 0]).  One warp per query row, a lane
// owns 8 of the 256 hidden channels (one 16-byte load), fp32 accumulation, shuffle reduction.  Replaces a bf16 -> fp32
// cast, an fp32 SIMT library GEMM (10 us for 800 x 256 x 4) and the elementwise kernel on the decoder's critical
// box -> pairing chain.
__global__ void __launch_bounds__(256)
box_head_refine_kernel(const __nv_bfloat16* __restrict__ hidden, int ldh, const float* __restrict__ W2,
                       const float* __restrict__ b2, const float* __restrict__ centers, float* __restrict__ boxes,
                       int M) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= M) return;
  const uint4 u = *reinterpret_cast<const uint4*>(hidden + static_cast<size_t>(r) * ldh + lane * 8);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  float h[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[2 * i] = __uint_as_float(w[i] << 16);
    h[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  float acc[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const float4 a = *reinterpret_cast<const float4*>(W2 + o * 256 + lane * 8);
    const float4 c = *reinterpret_cast<const float4*>(W2 + o * 256 + lane * 8 + 4);
    float t = h[0] * a.x;
    t = fmaf(h[1], a.y, t);
    t = fmaf(h[2], a.z, t);
    t = fmaf(h[3], a.w, t);
    t = fmaf(h[4], c.x, t);
    t = fmaf(h[5], c.y, t);
    t = fmaf(h[6], c.z, t);
    t = fmaf(h[7], c.w, t);
    acc[o] = t;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
  if (lane == 0) {
    const float2 c = reinterpret_cast<const float2*>(centers)[r];
    const float lx = -logf(__fsub_rn(__fdiv_rn(1.f, fmaxf(c.x, 1e-6f)), 1.f));
    const float ly = -logf(__fsub_rn(__fdiv_rn(1.f, fmaxf(c.y, 1e-6f)), 1.f));
    float4 o;
    o.x = 1.f / (1.f + expf(-(acc[0] + b2[0] + lx)));
    o.y = 1.f / (1.f + expf(-(acc[1] + b2[1] + ly)));
    o.z = 1.f / (1.f + expf(-(acc[2] + b2[2])));
    o.w = 1.f / (1.f + expf(-(acc[3] + b2[3])));
    reinterpret_cast<float4*>(boxes)[r] = o;
  }
}

// ---------------------------------------------------------------------------------------------
// matching cost, diagonal blocks only (matcher.py:72-107 / 158-184, bbox_utils.py:160-216)
// block = 8 predictions x 32 target lanes; grid = (ceil(Q/8), B)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
match_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                  const int32_t* __restrict__ tgt_ids, const float* __restrict__ tgt_boxes,
                  const int32_t* __restrict__ tgt_offsets, float* __restrict__ cost, int Q, int C, float w_class,
                  float w_bbox, float w_ciou, int with_l1) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  const int t0 = tgt_offsets[b], T = tgt_offsets[b + 1] - t0;
  if (T <= 0) return;
  const float4 pb = reinterpret_cast<const float4*>(boxes)[(size_t)b * Q + q];  // cx, cy, h, w
  const XYXY p = to_xyxy(pb.x, pb.y, pb.z, pb.w);
  // from_xyxy_to_cxcyhw(pred_xyxy) (bbox_utils.py:82-103)
  const float pcx = clamp01(__fdiv_rn(__fadd_rn(p.x0, p.x1), 2.f));
  const float pcy = clamp01(__fdiv_rn(__fadd_rn(p.y0, p.y1), 2.f));
  const float ph = clamp01(__fsub_rn(p.y1, p.y0));
  const float pw = clamp01(__fsub_rn(p.x1, p.x0));
  const float p_area = __fmul_rn(__fsub_rn(p.x1, p.x0), __fsub_rn(p.y1, p.y0));
  const float p_atan = atanf(__fdiv_rn(pw, fmaxf(ph, 1e-6f)));
  const float* lrow = logits + ((size_t)b * Q + q) * C;
  float* crow = cost + (size_t)Q * t0 + (size_t)q * T;
  const float four_over_pi2 = (float)(4.0 / (3.141592653589793 * 3.141592653589793));  // python double 4/pi**2 -> fp32
  for (int t = lane; t < T; t += 32) {
    const float4 g = reinterpret_cast<const float4*>(tgt_boxes)[t0 + t];  // x0,y0,x1,y1
    // ---- focal class cost (matcher.py:87-93) ----
    // a label outside [0, C) is an IndexError in the reference (matcher.py:91 `[:, tgt_ids]`); here its cost is NaN, which
    // the assignment kernel reports through its status flag (the matcher then raises, like scipy on NaN costs)
    const int cls_id = tgt_ids[t0 + t];
    const float lg = (cls_id >= 0 && cls_id < C) ? lrow[cls_id] : __int_as_float(0x7fc00000);
    const float prob = 1.f / (1.f + expf(-lg));
    const float om = __fsub_rn(1.f, prob);
    const float neg = __fmul_rn(__fmul_rn(0.75f, __fmul_rn(prob, prob)), -logf(__fadd_rn(om, 1e-8f)));
    const float pos = __fmul_rn(__fmul_rn(0.25f, __fmul_rn(om, om)), -logf(__fadd_rn(prob, 1e-8f)));
    const float c_cls = __fsub_rn(pos, neg);
    // ---- IoU (bbox_utils.py:201-216) ----
    const float iw = fmaxf(__fsub_rn(fminf(p.x1, g.z), fmaxf(p.x0, g.x)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(p.y1, g.w), fmaxf(p.y0, g.y)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float g_area = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
    const float uni = __fsub_rn(__fadd_rn(p_area, g_area), inter);
    const float iou = __fdiv_rn(inter, fmaxf(uni, 1e-6f));
    // ---- CIoU (bbox_utils.py:160-198) ----
    const float gcx = clamp01(__fdiv_rn(__fadd_rn(g.x, g.z), 2.f));
    const float gcy = clamp01(__fdiv_rn(__fadd_rn(g.y, g.w), 2.f));
    const float gh = clamp01(__fsub_rn(g.w, g.y));
    const float gw = clamp01(__fsub_rn(g.z, g.x));
    const float hw_ = fmaxf(__fsub_rn(fmaxf(p.x1, g.z), fminf(p.x0, g.x)), 0.f);
    const float hh_ = fmaxf(__fsub_rn(fmaxf(p.y1, g.w), fminf(p.y0, g.y)), 0.f);
    const float diag = __fadd_rn(__fmul_rn(hw_, hw_), __fmul_rn(hh_, hh_));
    const float dx = fabsf(__fsub_rn(pcx, gcx)), dy = fabsf(__fsub_rn(pcy, gcy));
    const float cdist = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float da = __fsub_rn(atanf(__fdiv_rn(gw, fmaxf(gh, 1e-6f))), p_atan);
    const float v = __fmul_rn(four_over_pi2, __fmul_rn(da, da));
    const float alpha = __fmul_rn(iou > 0.5f ? 1.f : 0.f, __fdiv_rn(v, __fadd_rn(__fsub_rn(1.f, iou), v)));
    const float raw = __fsub_rn(__fsub_rn(iou, __fdiv_rn(cdist, fmaxf(diag, 1e-6f))), __fmul_rn(alpha, v));
    const float ciou = isnan(raw) ? raw : fminf(fmaxf(raw, -1.f), 1.f);  // torch.clamp keeps NaN (identical boxes)
    const float c_iou = __fsub_rn(1.f, ciou);
    float total;
    if (with_l1) {
      // torch.cdist(p=1) of pred cxcyhw against target xyxy (matcher.py:96)
      float l1 = fabsf(__fsub_rn(pb.x, g.x));
      l1 = __fadd_rn(l1, fabsf(__fsub_rn(pb.y, g.y)));
      l1 = __fadd_rn(l1, fabsf(__fsub_rn(pb.z, g.z)));
      l1 = __fadd_rn(l1, fabsf(__fsub_rn(pb.w, g.w)));
      total = __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_class, c_cls)), __fmul_rn(w_ciou, c_iou));
    } else {
      total = __fadd_rn(__fmul_rn(w_class, c_cls), __fmul_rn(w_ciou, c_iou));
    }
    crow[t] = total;
  }
}

}  // namespace
}  // namespace destr

using namespace destr;

extern "C" int destr_pair_indices(const float* coords, int32_t* pairs, int B, int Q, void* stream) {
  DESTR_CHECK_ARG(coords && pairs && B > 0 && Q > 0 && Q <= 4096, "shape");
  dim3 grid(ceil_div(Q, 8), B);
  DESTR_SMEM_OPTIN(pair_indices_kernel, (size_t)Q * 6 * sizeof(float));  // Q > 2048 needs more than the default 48 KB
  pair_indices_kernel<<<grid, 256, (size_t)Q * 6 * sizeof(float), (cudaStream_t)stream>>>(coords, pairs, Q);
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_box_refine(const float* delta, const float* centers, float* boxes, int M, void* stream) {
  DESTR_CHECK_ARG(delta && centers && boxes && M > 0, "shape");
  box_refine_kernel<<<ceil_div(M, 128), 128, 0, (cudaStream_t)stream>>>(delta, centers, boxes, M);
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_box_head_refine(const void* hidden, int ldh, const float* W2, const float* b2, const float* centers,
                                     float* boxes, int M, void* stream) {
  DESTR_CHECK_ARG(hidden && W2 && b2 && centers && boxes && M > 0 && ldh >= 256 && ldh % 8 == 0, "shape");
  box_head_refine_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16*>(hidden), ldh,
                                                                          W2, b2, centers, boxes, M);
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_match_cost_blockdiag(const float* logits, const float* boxes, const int32_t* tgt_ids,
                                          const float* tgt_boxes, const int32_t* tgt_offsets, float* cost, int B,
                                          int Q, int C, float w_class, float w_bbox, float w_ciou, int with_l1,
                                          void* stream) {
  DESTR_CHECK_ARG(logits && boxes && tgt_ids && tgt_boxes && tgt_offsets && cost, "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && C > 0, "shape");
  dim3 grid(ceil_div(Q, 8), B);
  match_cost_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, boxes, tgt_ids, tgt_boxes, tgt_offsets, cost, Q,
                                                            C, w_class, w_bbox, w_ciou, with_l1);
  DESTR_LAUNCH_CHECK();
  return 0;
}
