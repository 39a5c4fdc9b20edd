// Query selection of the mini-detector (reference src/model/blocks/mini_detector.py:70-104, 142-170): per image the
// top-k positions by max-class score, the padding fix-up of get_topk_index, and the gathers of the selected
// object features (cls | reg) and box centres -- two small launches, no Python loop over the batch, no host round trip.
//
// Ordering key m[n] = max_c scores[n][c]: sigmoid is monotone, so this orders like the reference's
// max_c sigmoid(scores) and, where different m collapse to one fp32 sigmoid (the reference's order there is
// torch.topk's unspecified tie order), breaks the tie by position -- a refinement that needs no transcendental and is
// therefore bit-exact against the CPU oracle (oracle/query_select_oracle.py).
// Kernel 1 (keys): warp per position, lanes over the classes (coalesced), one redux.max on the order-preserving
// uint32 image of the float; also counts the un-padded positions of each image.
// Kernel 2 (rank + gather): top-k by ranking, rank[n] = #{j : m[j] > m[n] or (m[j] == m[n] and j < n)} -- no sort
// network, no cross-CTA dependency: the positions of an image are split over S CTAs (so that B*S fills the GPU), G
// lanes share the N comparisons of one position, and a position of rank r writes its own output rows: slot r and,
// for the padding fix-up (:86-98: with valid < k, slot s >= valid takes idx[valid - 1 - (s % valid)]), the slots
// s = (valid - 1 - r) + j*valid, j >= 1.
#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

__device__ __forceinline__ uint32_t ordered_u32(float v) {  // monotone float -> uint32 (-0 and +0 coincide)
  const uint32_t b = __float_as_uint(v + 0.0f);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

__global__ void __launch_bounds__(256)
query_keys_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ mask, int B, int N, int C,
                  uint32_t* __restrict__ keys, int32_t* __restrict__ valid_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t pos = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);  // b*N + n
  if (pos >= static_cast<int64_t>(B) * N) return;
  float v = -INFINITY;
  for (int c = lane; c < C; c += 32) v = fmaxf(v, scores[pos * C + c]);
  const uint32_t u = __reduce_max_sync(0xffffffffu, ordered_u32(v));
  if (lane == 0) {
    keys[pos] = u;
    if (!(mask && mask[pos])) atomicAdd(valid_cnt + pos / N, 1);
  }
}

__global__ void __launch_bounds__(1024)
select_queries_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ valid_cnt,
                      const float* __restrict__ cls_feat, const float* __restrict__ reg_feat,
                      const float* __restrict__ coords, int N, int D, int k, int per, int G,
                      int64_t* __restrict__ topk_idx, float* __restrict__ sel_f32, __nv_bfloat16* __restrict__ sel_bf16,
                      float* __restrict__ centers, int32_t* __restrict__ status) {
  extern __shared__ uint32_t sm_u[];
  uint32_t* key = sm_u;                                  // [N]
  int2* list = reinterpret_cast<int2*>(key + N + (N & 1));  // [k] (slot, position) pairs this CTA writes
  __shared__ int s_cnt;
  const int b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const int valid = valid_cnt[b];
  if (valid <= 0) {  // the reference divides by `valid` (:93): report instead
    if (blockIdx.x == 0 && tid == 0) status[b] = 1;
    return;
  }
  if (blockIdx.x == 0 && tid == 0) status[b] = 0;
  if (tid == 0) s_cnt = 0;
  for (int n = tid; n < N; n += nt) key[n] = keys[static_cast<size_t>(b) * N + n];
  __syncthreads();
  const int n0 = blockIdx.x * per, n1 = min(N, n0 + per);
  const int e = tid / G, g = tid - e * G, epp = nt / G;  // G (power of two <= 32) lanes per position
  const int top = min(k, valid);
  for (int nb = n0; nb < n1; nb += epp) {  // (uniform trip count: the shuffles below need whole warps)
    const int n = nb + e;
    const bool active = n < n1;
    int cnt = 0;
    if (active) {
      const uint32_t kn = key[n];
      for (int j = g; j < N; j += G) {
        const uint32_t kj = key[j];
        cnt += (kj > kn) || (kj == kn && j < n);
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (active && g == 0 && cnt < top) {
      list[atomicAdd(&s_cnt, 1)] = make_int2(cnt, n);
      for (int sl = (valid - 1 - cnt) + valid; sl < k; sl += valid) list[atomicAdd(&s_cnt, 1)] = make_int2(sl, n);
    }
  }
  __syncthreads();
  // rows of the listed slots: 4 floats per thread per step over [cls | reg]
  const int d4 = D / 4, row4 = 2 * d4, items = s_cnt * row4;
  for (int i = tid; i < items; i += nt) {
    const int li = i / row4, c4 = i - li * row4;
    const int sl = list[li].x, n = list[li].y;
    const size_t orow = static_cast<size_t>(b) * k + sl;
    if (c4 == 0) {
      topk_idx[orow] = n;
      centers[orow * 2] = coords[(static_cast<size_t>(b) * N + n) * 4];
      centers[orow * 2 + 1] = coords[(static_cast<size_t>(b) * N + n) * 4 + 1];
    }
    const float* src = (c4 < d4 ? cls_feat : reg_feat) + (static_cast<size_t>(b) * N + n) * D + (c4 % d4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src);
    const size_t o = orow * 2 * D + c4 * 4;
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
                                    int32_t* status, uint32_t* key_ws, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(scores && cls_feat && reg_feat && coords && topk_idx && centers && status && key_ws, "null pointer");
  DESTR_CHECK_ARG(sel_f32 || sel_bf16, "at least one of sel_f32 / sel_bf16");
  DESTR_CHECK_ARG(B > 0 && N > 0 && C > 0 && D > 0 && D % 4 == 0 && k > 0 && k <= N, "shape (D % 4 == 0, 1 <= k <= N)");
  const size_t smem = (static_cast<size_t>(N) + 1) * 4 + static_cast<size_t>(k) * 8;
  DESTR_CHECK_ARG(smem <= 200 * 1024, "N too large for the shared-memory ranking");
  DESTR_SMEM_OPTIN(select_queries_kernel, smem);
  cudaStream_t st = (cudaStream_t)stream;
  // key_ws: B*N keys followed by B counters of valid positions
  int32_t* valid_cnt = reinterpret_cast<int32_t*>(key_ws + static_cast<size_t>(B) * N);
  DESTR_CUDA(cudaMemsetAsync(valid_cnt, 0, sizeof(int32_t) * B, st));
  const int64_t positions = static_cast<int64_t>(B) * N;
  query_keys_kernel<<<static_cast<int>((positions + 7) / 8), 256, 0, st>>>(scores, mask, B, N, C, key_ws, valid_cnt);
  DESTR_LAUNCH_CHECK();
  // split every image over S CTAs so that B*S covers the GPU twice; G lanes share one position's N comparisons
  int S = ceil_div(2 * 148, B);
  if (S > ceil_div(N, 32)) S = ceil_div(N, 32);
  if (S < 1) S = 1;
  const int per = ceil_div(N, S);
  S = ceil_div(N, per);
  int G = 32;
  while (G > 1 && per * G > 1024) G >>= 1;
  select_queries_kernel<<<dim3(S, B), 1024, smem, st>>>(key_ws, valid_cnt, cls_feat, reg_feat, coords, N, D, k, per, G,
                                                       topk_idx, sel_f32, static_cast<__nv_bfloat16*>(sel_bf16),
                                                       centers, status);
  DESTR_LAUNCH_CHECK();
  return 0;
}
