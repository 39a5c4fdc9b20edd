// Prediction heads of the detector (reference model.py:120-131): class logits = Linear(256, C) on the class stream of
// the decoder output, boxes = sigmoid(MLP(256 -> 256 -> 4)(box stream) + [inverse_sigmoid(centers), 0, 0]).
// M = B*Q rows only (800 at the benchmark shape), fp32 weights and arithmetic as in the reference: far too small for
// tensor cores to matter -- what the step paid for was ~50 launches of SIMT sgemm / elementwise / reduce kernels
// (~0.3 ms of a 5 ms step).  Three launches here: forward, backward w.r.t. the rows, backward w.r.t. the weights.
//   dec      bf16 [M, 512]   decoder output, columns [0,256) class stream, [256,512) box stream
//   hidden   fp32 [M, 256]   relu(W1 x_box + b1), kept for backward
// Forward: a warp owns an output feature, lanes split the 256-long dot product (coalesced weight rows), 8 rows of the
// CTA share every weight load; the 8 partial sums are reduced with a 7-shuffle transpose-reduction.
// Backward rows: thread = input feature k (W[t][k] is contiguous in k), dY rows broadcast from shared memory.
// Backward weights: CTA = 4 output features, the M rows split over its 8 warps, lane = 8 input features.
#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

constexpr int HD = 256;     // hidden_dim
constexpr int HROWS = 8;    // rows per CTA (forward / backward-rows)
constexpr int HMAXC = 128;  // classes

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// sum of v[r] over the 32 lanes for r = 0..7; returned to the lane whose bits 4,3,2 spell r (lanes with lane&3 == 0)
__device__ __forceinline__ float reduce8(float (&v)[8], int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = (lane & 16) ? v[i] : v[i + 4], keep = (lane & 16) ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (lane & 8) ? v[i] : v[i + 2], keep = (lane & 8) ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float send = (lane & 4) ? v[0] : v[1], keep = (lane & 4) ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

// dot products of one weight row (256 fp32, this lane's 8 entries in w0, w1) with 8 rows of x (shared, fp32, pitch ldx)
__device__ __forceinline__ void dot8(const float4 w0, const float4 w1, const float* xs, int ldx, int lane, float (&acc)[8]) {
#pragma unroll
  for (int r = 0; r < HROWS; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(xs + r * ldx + lane * 8);
    const float4 b = *reinterpret_cast<const float4*>(xs + r * ldx + lane * 8 + 4);
    acc[r] = w0.x * a.x + w0.y * a.y + w0.z * a.z + w0.w * a.w + w1.x * b.x + w1.y * b.y + w1.z * b.z + w1.w * b.w;
  }
}

constexpr int FWD_THREADS = 512, FWD_WARPS = FWD_THREADS / 32;
__global__ void __launch_bounds__(FWD_THREADS)
heads_fwd_kernel(const __nv_bfloat16* __restrict__ dec, const float* __restrict__ centers, const float* __restrict__ Wc,
                 const float* __restrict__ bc, int C, const float* __restrict__ W1, const float* __restrict__ b1,
                 const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ logits,
                 float* __restrict__ boxes, float* __restrict__ hidden, int M) {
  __shared__ __align__(16) float xs[HROWS][2 * HD];  // the CTA's rows of dec, fp32
  __shared__ __align__(16) float hs[HROWS][HD];      // hidden activations
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * HROWS;
  for (int i = tid; i < HROWS * 2 * HD / 8; i += FWD_THREADS) {  // 8 bf16 per thread per step
    const int r = i / (2 * HD / 8), c = (i - r * (2 * HD / 8)) * 8;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (row0 + r < M) w = *reinterpret_cast<const uint4*>(dec + static_cast<size_t>(row0 + r) * 2 * HD + c);
    float* d = &xs[r][c];
    d[0] = bf_lo(w.x); d[1] = bf_hi(w.x); d[2] = bf_lo(w.y); d[3] = bf_hi(w.y);
    d[4] = bf_lo(w.z); d[5] = bf_hi(w.z); d[6] = bf_lo(w.w); d[7] = bf_hi(w.w);
  }
  __syncthreads();
  const int rsel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);  // row this lane ends up holding
  const bool writer = (lane & 3) == 0 && row0 + rsel < M;
  // this warp's outputs: hidden units warp, warp+16, .. then classes warp, warp+16, ..; the next weight row is
  // fetched while the current one is being used (the kernel is L2-latency bound, not throughput bound)
  auto wrow = [&](int o) -> const float* {
    return o < HD ? W1 + static_cast<size_t>(o) * HD : Wc + static_cast<size_t>(min(o - HD, C - 1)) * HD;
  };
  float4 n0 = *reinterpret_cast<const float4*>(wrow(warp) + lane * 8), n1 = *reinterpret_cast<const float4*>(wrow(warp) + lane * 8 + 4);
  // hidden layer of the box MLP
  for (int t = warp; t < HD; t += FWD_WARPS) {
    float acc[8];
    const float4 w0 = n0, w1 = n1;
    {
      const int nx = t + FWD_WARPS < HD ? t + FWD_WARPS : HD + warp;
      n0 = *reinterpret_cast<const float4*>(wrow(nx) + lane * 8);
      n1 = *reinterpret_cast<const float4*>(wrow(nx) + lane * 8 + 4);
    }
    dot8(w0, w1, &xs[0][HD], 2 * HD, lane, acc);
    const float s = reduce8(acc, lane);
    if ((lane & 3) == 0) {
      const float h = fmaxf(s + b1[t], 0.f);
      hs[rsel][t] = h;
      if (writer) hidden[static_cast<size_t>(row0 + rsel) * HD + t] = h;
    }
  }
  // class logits
  for (int c = warp; c < C; c += FWD_WARPS) {
    float acc[8];
    const float4 w0 = n0, w1 = n1;
    n0 = *reinterpret_cast<const float4*>(wrow(HD + c + FWD_WARPS) + lane * 8);
    n1 = *reinterpret_cast<const float4*>(wrow(HD + c + FWD_WARPS) + lane * 8 + 4);
    dot8(w0, w1, &xs[0][0], 2 * HD, lane, acc);
    const float s = reduce8(acc, lane);
    if (writer) logits[static_cast<size_t>(row0 + rsel) * C + c] = s + bc[c];
  }
  __syncthreads();
  // box output layer + reference point + sigmoid
  if (warp < 4) {
    const int j = warp;
    float acc[8];
    dot8(*reinterpret_cast<const float4*>(W2 + static_cast<size_t>(j) * HD + lane * 8),
         *reinterpret_cast<const float4*>(W2 + static_cast<size_t>(j) * HD + lane * 8 + 4), &hs[0][0], HD, lane, acc);
    const float s = reduce8(acc, lane);
    if (writer) {
      float z = s + b2[j];
      if (j < 2) {  // + inverse_sigmoid(center)  (misc.py:59-62: -log(1 / max(x, 1e-6) - 1))
        const float x = fmaxf(centers[static_cast<size_t>(row0 + rsel) * 2 + j], 1e-6f);
        z += -logf(1.f / x - 1.f);
      }
      boxes[static_cast<size_t>(row0 + rsel) * 4 + j] = 1.f / (1.f + expf(-z));
    }
  }
}

// gradient w.r.t. the rows: d_dec = [dlogits Wc | (dz W2 o relu') W1] in bf16; also leaves dh and dz for the weight pass
__global__ void __launch_bounds__(512)
heads_bwd_rows_kernel(const float* __restrict__ hidden, const float* __restrict__ boxes, const float* __restrict__ dlogits,
                      const float* __restrict__ dboxes, const float* __restrict__ Wc, const float* __restrict__ W1,
                      const float* __restrict__ W2, int C, __nv_bfloat16* __restrict__ d_dec, float* __restrict__ dh_ws,
                      float* __restrict__ dz_ws, int M) {
  __shared__ __align__(16) float dl[HMAXC][HROWS];  // dlogits, transposed: 8 rows of one class are contiguous
  __shared__ __align__(16) float dh[HD][HROWS];
  __shared__ float dz[HROWS][4];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * HROWS;
  for (int i = tid; i < HROWS * C; i += 512) {
    const int r = i / C, c = i - r * C;
    dl[c][r] = row0 + r < M ? dlogits[static_cast<size_t>(row0 + r) * C + c] : 0.f;
  }
  if (tid < HROWS * 4) {
    const int r = tid >> 2, j = tid & 3;
    float v = 0.f;
    if (row0 + r < M) {
      const float s = boxes[static_cast<size_t>(row0 + r) * 4 + j];
      v = dboxes[static_cast<size_t>(row0 + r) * 4 + j] * s * (1.f - s);  // through the sigmoid
      dz_ws[static_cast<size_t>(row0 + r) * 4 + j] = v;
    }
    dz[r][j] = v;
  }
  __syncthreads();
  if (tid < HD) {  // dh[r][t] = relu'(h) * sum_j dz[r][j] W2[j][t]      (thread = t)
    const int t = tid;
    const float w0 = W2[t], w1 = W2[HD + t], w2 = W2[2 * HD + t], w3 = W2[3 * HD + t];
#pragma unroll
    for (int r = 0; r < HROWS; ++r) {
      float v = 0.f;
      if (row0 + r < M) {
        const float h = hidden[static_cast<size_t>(row0 + r) * HD + t];
        v = h > 0.f ? dz[r][0] * w0 + dz[r][1] * w1 + dz[r][2] * w2 + dz[r][3] * w3 : 0.f;
        dh_ws[static_cast<size_t>(row0 + r) * HD + t] = v;
      }
      dh[t][r] = v;
    }
  }
  __syncthreads();
  // threads [0,256): box stream, d = dh W1;  threads [256,512): class stream, d = dlogits Wc     (thread = k)
  const int k = tid & (HD - 1);
  const bool box = tid < HD;
  const float* W = box ? W1 : Wc;
  const float(*gsm)[HROWS] = box ? dh : dl;
  const int n = box ? HD : C;
  float a[HROWS];
#pragma unroll
  for (int r = 0; r < HROWS; ++r) a[r] = 0.f;
#pragma unroll 16
  for (int t = 0; t < n; ++t) {
    const float w = W[static_cast<size_t>(t) * HD + k];
    const float4 g0 = *reinterpret_cast<const float4*>(&gsm[t][0]), g1 = *reinterpret_cast<const float4*>(&gsm[t][4]);
    a[0] += g0.x * w; a[1] += g0.y * w; a[2] += g0.z * w; a[3] += g0.w * w;
    a[4] += g1.x * w; a[5] += g1.y * w; a[6] += g1.z * w; a[7] += g1.w * w;
  }
#pragma unroll
  for (int r = 0; r < HROWS; ++r)
    if (row0 + r < M) d_dec[static_cast<size_t>(row0 + r) * 2 * HD + (box ? HD : 0) + k] = __float2bfloat16(a[r]);
}

// gradient w.r.t. the weights: dW[o][k] = sum_r G[r][o] X[r][k], db[o] = sum_r G[r][o] for the three layers.
//   CTA = 4 output features o of one layer; warp w takes rows r = w, w+8, ..; lane = 8 consecutive input features k;
//   the 8 warps' partial sums are added in a fixed order through shared memory (deterministic, no atomics)
constexpr int WB = 4;
__global__ void __launch_bounds__(256)
heads_bwd_weights_kernel(const __nv_bfloat16* __restrict__ dec, const float* __restrict__ hidden,
                         const float* __restrict__ dlogits, const float* __restrict__ dh_ws,
                         const float* __restrict__ dz_ws, int C, float* __restrict__ dWc, float* __restrict__ dbc,
                         float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dW2,
                         float* __restrict__ db2, int M) {
  __shared__ __align__(16) float red[8][WB][HD];
  __shared__ float redb[8][WB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n1 = HD / WB, nc = (C + WB - 1) / WB;
  int blk = blockIdx.x;
  const float* G;
  int ldg, o0, n_o, layer;
  float *dW, *db;
  if (blk < n1) {
    layer = 0; G = dh_ws; ldg = HD; o0 = blk * WB; n_o = WB; dW = dW1; db = db1;
  } else if (blk < n1 + nc) {
    layer = 1; blk -= n1; G = dlogits; ldg = C; o0 = blk * WB; n_o = min(WB, C - o0); dW = dWc; db = dbc;
  } else {
    layer = 2; G = dz_ws; ldg = 4; o0 = 0; n_o = 4; dW = dW2; db = db2;
  }
  float acc[WB][8], gs[WB];
#pragma unroll
  for (int i = 0; i < WB; ++i) {
    gs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }
  // X[r][k]: box stream (layer 0) / class stream (layer 1) of dec, hidden (layer 2)
  const int xcol = (layer == 0 ? HD : 0) + lane * 8;
#pragma unroll 4
  for (int r = warp; r < M; r += 8) {
    float x[8];
    if (layer == 2) {
      const float4 a = *reinterpret_cast<const float4*>(hidden + static_cast<size_t>(r) * HD + lane * 8);
      const float4 b = *reinterpret_cast<const float4*>(hidden + static_cast<size_t>(r) * HD + lane * 8 + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
      const uint4 w = *reinterpret_cast<const uint4*>(dec + static_cast<size_t>(r) * 2 * HD + xcol);
      x[0] = bf_lo(w.x); x[1] = bf_hi(w.x); x[2] = bf_lo(w.y); x[3] = bf_hi(w.y);
      x[4] = bf_lo(w.z); x[5] = bf_hi(w.z); x[6] = bf_lo(w.w); x[7] = bf_hi(w.w);
    }
    const float* g = G + static_cast<size_t>(r) * ldg + o0;
#pragma unroll
    for (int i = 0; i < WB; ++i) {
      const float gi = i < n_o ? g[i] : 0.f;
      gs[i] += gi;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] += gi * x[j];
    }
  }
#pragma unroll
  for (int i = 0; i < WB; ++i) {
    *reinterpret_cast<float4*>(&red[warp][i][lane * 8]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(&red[warp][i][lane * 8 + 4]) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    if (lane == 0) redb[warp][i] = gs[i];
  }
  __syncthreads();
  for (int i = 0; i < n_o; ++i) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][i][tid];
    dW[static_cast<size_t>(o0 + i) * HD + tid] = sum;
  }
  if (tid < n_o) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += redb[w][tid];
    db[o0 + tid] = sum;
  }
}

}  // namespace
}  // namespace destr

using namespace destr;

extern "C" int destr_heads_fwd(const void* dec, const float* centers, const float* Wc, const float* bc, int C,
                               const float* W1, const float* b1, const float* W2, const float* b2, float* logits,
                               float* boxes, float* hidden, int M, void* stream) {
  DESTR_CHECK_ARG(dec && centers && Wc && bc && W1 && b1 && W2 && b2 && logits && boxes && hidden, "null pointer");
  DESTR_CHECK_ARG(M > 0 && C > 0 && C <= HMAXC, "shape (1 <= C <= 128)");
  heads_fwd_kernel<<<ceil_div(M, HROWS), FWD_THREADS, 0, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(dec), centers, Wc, bc, C, W1, b1, W2, b2, logits, boxes, hidden, M);
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_heads_bwd(const void* dec, const float* hidden, const float* boxes, const float* dlogits,
                               const float* dboxes, const float* Wc, const float* W1, const float* W2, int C,
                               void* d_dec, float* dh_ws, float* dz_ws, float* dWc, float* dbc, float* dW1,
                               float* db1, float* dW2, float* db2, int M, int which, void* stream) {
  DESTR_CHECK_ARG(dec && hidden && boxes && dlogits && dboxes && Wc && W1 && W2 && d_dec && dh_ws && dz_ws,
                  "null pointer");
  DESTR_CHECK_ARG(which >= 1 && which <= 3, "which: 1 = row gradients (d_dec, workspaces), 2 = parameter gradients, 3 = both");
  DESTR_CHECK_ARG(dWc && dbc && dW1 && db1 && dW2 && db2, "null gradient pointer");
  DESTR_CHECK_ARG(M > 0 && C > 0 && C <= HMAXC, "shape (1 <= C <= 128)");
  cudaStream_t st = (cudaStream_t)stream;
  if (which & 1) {
    heads_bwd_rows_kernel<<<ceil_div(M, HROWS), 512, 0, st>>>(hidden, boxes, dlogits, dboxes, Wc, W1, W2, C,
                                                            static_cast<__nv_bfloat16*>(d_dec), dh_ws, dz_ws, M);
    DESTR_LAUNCH_CHECK();
  }
  if (which & 2) {  // reads the dh / dz workspaces of the row kernel; nothing downstream of d_dec waits for it
    const int blocks = HD / WB + ceil_div(C, WB) + 1;
    heads_bwd_weights_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dec), hidden, dlogits, dh_ws,
                                                     dz_ws, C, dWc, dbc, dW1, db1, dW2, db2, M);
    DESTR_LAUNCH_CHECK();
  }
  return 0;
}
