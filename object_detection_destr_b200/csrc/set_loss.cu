// Fused set-prediction loss, forward AND backward in one launch.
//
// Replaces the per-image Python loop of SetCriterion.forward after matching
// (reference src/utils/criterion.py:29-79 with loss_fn = {class: sigmoid_focal_loss (src/utils/misc.py:99-128),
// bbox: L1Loss, ciou: CompleteIOULoss (criterion.py:82-89: the MEAN OF THE FULL n x n complete_iou matrix,
// src/utils/bbox_utils.py:160-198)}) and its autograd graph -- ~350 tiny elementwise launches per step -- by
// one kernel that produces the three losses, their weighted total, and d(total)/d(logits), d(total)/d(boxes).
//
// One CTA per image.  Gradients follow torch's autograd conventions exactly where they are not the textbook
// derivative: clamp passes the gradient on the closed interval, maximum/minimum split it in half on ties,
// abs uses sign() (0 at 0), alpha of the CIoU aspect term carries no gradient (bbox_utils.py:191-193).
// The box path is evaluated with forward-mode dual numbers (value + 4 partials w.r.t. the predicted box),
// so the formula is written once, in the reference's operation order.
// The cross-image reduction is done by the last CTA to finish, in image order (deterministic).
#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

struct Dual {  // value and gradient w.r.t. the predicted (cx, cy, h, w)
  float v, d[4];
};
__device__ __forceinline__ Dual dconst(float v) { return Dual{v, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ Dual dvar(float v, int i) {
  Dual r = dconst(v);
  r.d[i] = 1.f;
  return r;
}
__device__ __forceinline__ Dual operator+(const Dual& a, const Dual& b) {
  Dual r{a.v + b.v, {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
__device__ __forceinline__ Dual operator-(const Dual& a, const Dual& b) {
  Dual r{a.v - b.v, {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
__device__ __forceinline__ Dual operator*(const Dual& a, const Dual& b) {
  Dual r{a.v * b.v, {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
__device__ __forceinline__ Dual operator/(const Dual& a, const Dual& b) {
  const float q = a.v / b.v;
  Dual r{q, {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = (a.d[i] - q * b.d[i]) / b.v;
  return r;
}
__device__ __forceinline__ Dual scale(const Dual& a, float s) {
  Dual r{a.v * s, {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * s;
  return r;
}
// torch.clamp: gradient passes where lo <= x <= hi (closed interval)
__device__ __forceinline__ Dual dclamp(const Dual& a, float lo, float hi) {
  const bool pass = (a.v >= lo) && (a.v <= hi);
  Dual r{fminf(fmaxf(a.v, lo), hi), {}};
  if (isnan(a.v)) r.v = a.v;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = pass ? a.d[i] : 0.f;
  return r;
}
__device__ __forceinline__ Dual dclamp_min(const Dual& a, float lo) { return dclamp(a, lo, INFINITY); }
__device__ __forceinline__ Dual dclamp_max(const Dual& a, float hi) { return dclamp(a, -INFINITY, hi); }
// torch.maximum / minimum against a constant: ties give half the gradient
__device__ __forceinline__ Dual dmaxc(const Dual& a, float c) {
  const float w = (a.v > c) ? 1.f : ((a.v == c) ? 0.5f : 0.f);
  Dual r{fmaxf(a.v, c), {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = w * a.d[i];
  return r;
}
__device__ __forceinline__ Dual dminc(const Dual& a, float c) {
  const float w = (a.v < c) ? 1.f : ((a.v == c) ? 0.5f : 0.f);
  Dual r{fminf(a.v, c), {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = w * a.d[i];
  return r;
}
__device__ __forceinline__ Dual dabs(const Dual& a) {
  const float s = (a.v > 0.f) ? 1.f : ((a.v < 0.f) ? -1.f : 0.f);
  Dual r{fabsf(a.v), {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = s * a.d[i];
  return r;
}
__device__ __forceinline__ Dual datan(const Dual& a) {
  const float g = 1.f / (1.f + a.v * a.v);
  Dual r{atanf(a.v), {}};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = g * a.d[i];
  return r;
}

struct PredBox {  // xyxy of a predicted cxcyhw box (from_cxcyhw_to_xyxy, bbox_utils.py:33-63) and derived terms
  Dual x0, y0, x1, y1, cx, cy, h, w, area, at;
};
__device__ PredBox make_pred(const float* b) {
  const Dual cx = dvar(b[0], 0), cy = dvar(b[1], 1), hh = dvar(b[2], 2), ww = dvar(b[3], 3);
  PredBox p;
  p.x0 = dclamp_min(cx - scale(ww, 0.5f), 0.f);
  p.y0 = dclamp_min(cy - scale(hh, 0.5f), 0.f);
  p.x1 = dclamp_max(cx + scale(ww, 0.5f), 1.f);
  p.y1 = dclamp_max(cy + scale(hh, 0.5f), 1.f);
  // from_xyxy_to_cxcyhw (bbox_utils.py:66-103): everything clipped to [0,1]
  p.cx = dclamp(scale(p.x0 + p.x1, 0.5f), 0.f, 1.f);
  p.cy = dclamp(scale(p.y0 + p.y1, 0.5f), 0.f, 1.f);
  p.h = dclamp(p.y1 - p.y0, 0.f, 1.f);
  p.w = dclamp(p.x1 - p.x0, 0.f, 1.f);
  p.area = (p.x1 - p.x0) * (p.y1 - p.y0);
  p.at = datan(p.w / dclamp_min(p.h, 1e-6f));
  return p;
}

// complete_iou cost of one (prediction, target) pair: 1 - clamp(IoU - rho^2/c^2 - alpha v, -1, 1)
__device__ Dual ciou_cost(const PredBox& p, const float* g) {
  const float eps = 1e-6f;
  const float gx0 = g[0], gy0 = g[1], gx1 = g[2], gy1 = g[3];
  const float gcx = fminf(fmaxf((gx0 + gx1) * 0.5f, 0.f), 1.f), gcy = fminf(fmaxf((gy0 + gy1) * 0.5f, 0.f), 1.f);
  const float gh = fminf(fmaxf(gy1 - gy0, 0.f), 1.f), gw = fminf(fmaxf(gx1 - gx0, 0.f), 1.f);
  const Dual iw = dclamp_min(dminc(p.x1, gx1) - dmaxc(p.x0, gx0), 0.f);
  const Dual ih = dclamp_min(dminc(p.y1, gy1) - dmaxc(p.y0, gy0), 0.f);
  const Dual inter = iw * ih;
  const float area_g = (gx1 - gx0) * (gy1 - gy0);
  const Dual iou = inter / dclamp_min(p.area + dconst(area_g) - inter, eps);
  const Dual hw = dclamp_min(dmaxc(p.x1, gx1) - dminc(p.x0, gx0), 0.f);
  const Dual hh = dclamp_min(dmaxc(p.y1, gy1) - dminc(p.y0, gy0), 0.f);
  const Dual c2 = hw * hw + hh * hh;
  const Dual dx = dabs(p.cx - dconst(gcx)), dy = dabs(p.cy - dconst(gcy));
  const Dual rho2 = dx * dx + dy * dy;
  const float at_g = atanf(gw / fmaxf(gh, eps));
  const Dual da = dconst(at_g) - p.at;
  const Dual v = scale(da * da, 4.0f / (3.14159265358979323846f * 3.14159265358979323846f));
  const float alpha = (iou.v > 0.5f ? 1.f : 0.f) * (v.v / (1.f - iou.v + v.v));  // no gradient (bbox_utils.py:191-193)
  const Dual ciou = dclamp(iou - rho2 / dclamp_min(c2, eps) - scale(v, alpha), -1.f, 1.f);
  return dconst(1.f) - ciou;
}

__device__ float block_sum(float v, float* red) {  // blockDim.x <= 1024, result valid in every thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

constexpr int kMaxQ = 1024;

__global__ void __launch_bounds__(512)
set_loss_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, const int64_t* __restrict__ tl,
                const float* __restrict__ tb, const int64_t* __restrict__ pi, const int64_t* __restrict__ ti,
                const uint8_t* __restrict__ valid, int B, int Q, int C, int Tm, int n, float w_class, float w_bbox,
                float w_ciou, float* __restrict__ losses, float* __restrict__ dlogits, float* __restrict__ dboxes,
                float* __restrict__ partial, unsigned int* __restrict__ counter) {
  __shared__ int cls_of_q[kMaxQ];
  __shared__ float red[32];
  __shared__ int s_cnt, s_images_with_targets;
  __shared__ bool s_last;
  const int b = blockIdx.x, tid = threadIdx.x;

  // ---- target class per query: "no object" = class 1 (criterion.py:41-44), matched queries get their label ----
  for (int q = tid; q < Q; q += blockDim.x) cls_of_q[q] = 1;
  if (tid == 0) {
    int cnt = 0;
    for (int k = 0; k < n; ++k) cnt += valid[b * n + k] ? 1 : 0;
    int has = 0;
    for (int bb = 0; bb < B; ++bb) {
      int any = 0;
      for (int k = 0; k < n; ++k) any |= valid[bb * n + k];
      has += any ? 1 : 0;
    }
    s_cnt = cnt;
    s_images_with_targets = has;
  }
  __syncthreads();
  for (int k = tid; k < n; k += blockDim.x)
    if (valid[b * n + k]) cls_of_q[pi[b * n + k]] = static_cast<int>(tl[b * Tm + ti[b * n + k]]);
  for (int i = tid; i < Q * 4; i += blockDim.x) dboxes[static_cast<size_t>(b) * Q * 4 + i] = 0.f;
  __syncthreads();

  // ---- sigmoid focal loss (misc.py:99-128), alpha = 0.25, gamma = 2; mean over classes, sum over queries / Q ----
  const float gscale_cls = w_class / (static_cast<float>(B) * Q * C);
  float acc_cls = 0.f;
  const float* lg = logits + static_cast<size_t>(b) * Q * C;
  float* dlg = dlogits + static_cast<size_t>(b) * Q * C;
  for (int i = tid; i < Q * C; i += blockDim.x) {
    const int q = i / C, c = i - q * C;
    const float x = lg[i];
    const bool t = (c == cls_of_q[q]);
    const float p = 1.f / (1.f + expf(-x));
    const float ce = fmaxf(x, 0.f) - (t ? x : 0.f) + log1pf(expf(-fabsf(x)));  // BCE with logits
    const float pt = t ? p : 1.f - p;
    const float a = t ? 0.25f : 0.75f;
    const float om = 1.f - pt;
    acc_cls += a * ce * om * om;
    const float dpt = t ? p * (1.f - p) : -p * (1.f - p);
    const float dce = p - (t ? 1.f : 0.f);
    dlg[i] = gscale_cls * a * (dce * om * om - 2.f * ce * om * dpt);
  }
  const float cls_img = block_sum(acc_cls, red) / (static_cast<float>(Q) * C);

  // ---- L1 (mean over the matched boxes' 4 coordinates) and CIoU (mean of the full cnt x cnt matrix) ----
  const int cnt = s_cnt;
  float acc_l1 = 0.f, acc_ci = 0.f;
  if (cnt > 0) {
    const float denom = static_cast<float>(s_images_with_targets);
    const float g_l1 = w_bbox / (4.f * cnt * denom), g_ci = w_ciou / (static_cast<float>(cnt) * cnt * denom);
    // one warp per matched prediction i, lanes over its targets j (long dependent dual-number chains: spreading
    // the cnt x cnt pairs over the block is what makes this kernel short); fixed-order shuffle reductions
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int k = warp; k < n; k += nwarps) {
      if (!valid[b * n + k]) continue;  // warp-uniform
      const int q = static_cast<int>(pi[b * n + k]);
      const float* pb = boxes + (static_cast<size_t>(b) * Q + q) * 4;
      const PredBox p = make_pred(pb);
      float grad[4] = {0.f, 0.f, 0.f, 0.f};
      float ci = 0.f;
      for (int j = lane; j < n; j += 32) {  // CIoU against EVERY matched target of the image (criterion.py:87-89)
        if (!valid[b * n + j]) continue;
        const Dual c = ciou_cost(p, tb + (static_cast<size_t>(b) * Tm + ti[b * n + j]) * 4);
        ci += c.v;
#pragma unroll
        for (int i = 0; i < 4; ++i) grad[i] += g_ci * c.d[i];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ci += __shfl_xor_sync(0xffffffffu, ci, o);
#pragma unroll
        for (int i = 0; i < 4; ++i) grad[i] += __shfl_xor_sync(0xffffffffu, grad[i], o);
      }
      if (lane == 0) {
        acc_ci += ci;
        // L1 against its own target
        const float* g = tb + (static_cast<size_t>(b) * Tm + ti[b * n + k]) * 4;
        const Dual e = dabs(p.x0 - dconst(g[0])) + dabs(p.y0 - dconst(g[1])) + dabs(p.x1 - dconst(g[2])) +
                       dabs(p.y1 - dconst(g[3]));
        acc_l1 += e.v;
        float* db = dboxes + (static_cast<size_t>(b) * Q + q) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) db[i] = grad[i] + g_l1 * e.d[i];
      }
    }
  }
  const float l1_img = block_sum(acc_l1, red), ci_img = block_sum(acc_ci, red);
  if (tid == 0) {
    partial[b * 3 + 0] = cls_img;
    partial[b * 3 + 1] = cnt > 0 ? l1_img / (4.f * cnt) : 0.f;
    partial[b * 3 + 2] = cnt > 0 ? ci_img / (static_cast<float>(cnt) * cnt) : 0.f;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == static_cast<unsigned int>(B) - 1u);
  }
  __syncthreads();
  if (s_last && tid == 0) {  // the last CTA reduces over images in index order
    __threadfence();
    float c = 0.f, l = 0.f, i = 0.f;
    for (int bb = 0; bb < B; ++bb) {
      c += partial[bb * 3 + 0];
      l += partial[bb * 3 + 1];
      i += partial[bb * 3 + 2];
    }
    const float denom = fmaxf(static_cast<float>(s_images_with_targets), 1.f);
    losses[0] = c / B;
    losses[1] = l / denom;
    losses[2] = i / denom;
    losses[3] = w_class * losses[0] + w_bbox * losses[1] + w_ciou * losses[2];
    *counter = 0u;
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_set_loss_fwd_bwd(const float* logits, const float* boxes, const int64_t* tgt_labels,
                                      const float* tgt_boxes, const int64_t* pred_idx, const int64_t* tgt_idx,
                                      const uint8_t* valid, int B, int Q, int C, int t_max, int n, float w_class,
                                      float w_bbox, float w_ciou, float* losses, float* dlogits, float* dboxes,
                                      float* workspace, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(logits && boxes && tgt_labels && tgt_boxes && pred_idx && tgt_idx && valid && losses && dlogits &&
                      dboxes && workspace, "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && Q <= kMaxQ && C > 0 && t_max > 0 && n > 0, "shape (Q <= 1024)");
  float* partial = workspace;
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace + 3 * B);
  set_loss_kernel<<<B, 512, 0, static_cast<cudaStream_t>(stream)>>>(logits, boxes, tgt_labels, tgt_boxes, pred_idx,
                                                                    tgt_idx, valid, B, Q, C, t_max, n, w_class,
                                                                    w_bbox, w_ciou, losses, dlogits, dboxes, partial,
                                                                    counter);
  DESTR_LAUNCH_CHECK();
  return 0;
}
