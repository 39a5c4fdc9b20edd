// Query selection of the mini-detector (reference src/model/blocks/mini_detector.py:70-104, 142-170): per image the
// top-k positions by max-class score, the padding fix-up of get_topk_index, and the gathers of the selected
// object features (cls | reg) and box centres -- one launch, no Python loop over the batch, no host round trip.
//
// One CTA per image.  Ordering key m[n] = max_c scores[n][c]: sigmoid is monotone, so this orders like the
// reference's max_c sigmoid(scores) and, where different m collapse to one fp32 sigmoid (the reference's order there
// is torch.topk's unspecified tie order), breaks the tie by position -- a refinement that needs no transcendental and
// is therefore bit-exact against the CPU oracle (oracle/query_select_oracle.py).
// Top-k by ranking: rank[n] = #{j : m[j] > m[n] or (m[j] == m[n] and j < n)}; the element of rank r < k goes to slot
// r.  N <= a few thousand keys sit in shared memory, so the N^2 comparisons cost ~N^2/1024 broadcast reads per thread
// (1.1 k at N = 1050) -- far below a sort's synchronisation cost at this size.
// Fix-up (:86-98): with valid = N - #padded < k, slot s >= valid takes idx[valid - 1 - (s % valid)].
#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

__global__ void __launch_bounds__(1024)
select_queries_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ mask,
                      const float* __restrict__ cls_feat, const float* __restrict__ reg_feat,
                      const float* __restrict__ coords, int N, int C, int D, int k, int64_t* __restrict__ topk_idx,
                      float* __restrict__ sel_f32, __nv_bfloat16* __restrict__ sel_bf16, float* __restrict__ centers,
                      int32_t* __restrict__ status) {
  extern __shared__ float sm_f[];
  float* m = sm_f;                                   // [N]
  int* sorted = reinterpret_cast<int*>(m + N);       // [k]  position of rank r
  int* fin = sorted + k;                             // [k]  after the fix-up
  __shared__ int s_pad;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* sc = scores + static_cast<size_t>(b) * N * C;
  if (tid == 0) s_pad = 0;
  __syncthreads();
  int pad = 0;
  for (int n = tid; n < N; n += nt) {
    float v = sc[static_cast<size_t>(n) * C];
    for (int c = 1; c < C; ++c) v = fmaxf(v, sc[static_cast<size_t>(n) * C + c]);
    m[n] = v;
    if (mask) pad += mask[static_cast<size_t>(b) * N + n] ? 1 : 0;
  }
  if (pad) atomicAdd(&s_pad, pad);
  __syncthreads();
  for (int n = tid; n < N; n += nt) {
    const float v = m[n];
    int rank = 0;
    for (int j = 0; j < N; ++j) {
      const float w = m[j];  // same address across the warp: one broadcast read
      rank += (w > v) || (w == v && j < n);
    }
    if (rank < k) sorted[rank] = n;
  }
  __syncthreads();
  const int valid = N - s_pad;
  if (valid <= 0) {  // the reference divides by `valid` (:93): report instead
    if (tid == 0) status[b] = 1;
    return;
  }
  if (tid == 0) status[b] = 0;
  for (int s = tid; s < k; s += nt) {
    const int src = (mask == nullptr || s < valid) ? s : valid - 1 - (s % valid);
    const int n = sorted[src];
    fin[s] = n;
    topk_idx[static_cast<size_t>(b) * k + s] = n;
    centers[(static_cast<size_t>(b) * k + s) * 2] = coords[(static_cast<size_t>(b) * N + n) * 4];
    centers[(static_cast<size_t>(b) * k + s) * 2 + 1] = coords[(static_cast<size_t>(b) * N + n) * 4 + 1];
  }
  __syncthreads();
  // gather [cls | reg] rows: 4 floats per thread per step
  const int d4 = D / 4, row4 = 2 * d4;
  for (int i = tid; i < k * row4; i += nt) {
    const int s = i / row4, c4 = i - s * row4;
    const int n = fin[s];
    const float* src = (c4 < d4 ? cls_feat : reg_feat) + (static_cast<size_t>(b) * N + n) * D + (c4 % d4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src);
    const size_t o = (static_cast<size_t>(b) * k + s) * 2 * D + c4 * 4;
    if (sel_f32) *reinterpret_cast<float4*>(sel_f32 + o) = v;
    if (sel_bf16) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 w;
      w.x = *reinterpret_cast<const uint32_t*>(&lo);
      w.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(sel_bf16 + o) = w;
    }
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_select_queries(const float* scores, const uint8_t* mask, const float* cls_feat,
                                    const float* reg_feat, const float* coords, int B, int N, int C, int D, int k,
                                    int64_t* topk_idx, float* sel_f32, void* sel_bf16, float* centers,
                                    int32_t* status, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(scores && cls_feat && reg_feat && coords && topk_idx && centers && status, "null pointer");
  DESTR_CHECK_ARG(sel_f32 || sel_bf16, "at least one of sel_f32 / sel_bf16");
  DESTR_CHECK_ARG(B > 0 && N > 0 && C > 0 && D > 0 && D % 4 == 0 && k > 0 && k <= N, "shape (D % 4 == 0, 1 <= k <= N)");
  const size_t smem = static_cast<size_t>(N) * 4 + static_cast<size_t>(k) * 8;
  DESTR_CHECK_ARG(smem <= 200 * 1024, "N too large for the shared-memory ranking");
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    DESTR_CUDA(cudaFuncSetAttribute(select_queries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  select_queries_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(scores, mask, cls_feat, reg_feat, coords, N, C, D, k,
                                                                topk_idx, sel_f32,
                                                                static_cast<__nv_bfloat16*>(sel_bf16), centers, status);
  DESTR_LAUNCH_CHECK();
  return 0;
}
