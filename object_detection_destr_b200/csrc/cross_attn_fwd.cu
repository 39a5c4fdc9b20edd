// Split cross-attention forward of both ClsRegBranch'es on tcgen05 / TMEM / TMA.
//
// Reference: DecoderBlock.forward src/model/blocks/decoder_block.py:189-217 builds q_cls/q_reg
// (512 = per-head interleave of q_obj-half and q_pos) and k (interleave of k_enc and k_pos), then
// ClsRegBranch (decoder_block.py:246-251) calls SelfAttention with ONE head: softmax(q.k^T/sqrt(512)
// + key-padding mask).v2.  The interleave is one fixed permutation of the 512 contraction index
// applied to both q and k, so  q.k = <q_obj_half, k_enc> + <q_pos, k_pos>  and no shuffle is needed.
//
// grid = (key tiles of 128, 2 branches x query tiles of 128, B).  Each CTA computes, for ONE key tile:
//   S = sum over 8 contraction chunks of 64  (4 from q_obj-half x k_enc, 4 from q_pos x k_pos)
//       streamed by TMA through a 4-stage ring (A and B chunk = 16 KB each, SW128, K-major)
//   m = rowmax, P = exp2(c*S - m) (bf16, written over S in TMEM), l = rowsum
//   O = P.V_tile  (V tile [128 keys][256] as four MN-major SW128 chunks; four N=64 MMAs)
// and stores the un-normalised partial (O fp32, m, l).  cross_attn_combine_kernel merges the key
// tiles:  out = sum_j 2^(m_j-M) O_j / sum_j 2^(m_j-M) l_j.   (split-KV: the decoder has only
// 2*ceil(Q/128)*B query tiles, far fewer than 148 SMs.)
// warps 0-7 softmax/epilogue (thread <-> query row x half of the key columns: warps w and w+4 own the same TMEM lanes and
// exchange the row maximum / row sum through shared memory), warp 8 TMA, warp 9 MMA.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {
namespace {

constexpr int BT = 128;
constexpr int NSTAGE = 4;
constexpr int NCHUNK = 8;   // 512 / 64
constexpr int DV = 256;
constexpr int NTHREADS = 320;  // 8 math warps (warp pairs share TMEM lanes, split the key columns) + TMA + MMA
constexpr uint32_t CHUNK_BYTES = BT * 128;  // 16 KB

struct __align__(1024) Smem {
  uint8_t a[NSTAGE][CHUNK_BYTES];
  uint8_t b[NSTAGE][CHUNK_BYTES];
  uint8_t v[DV / 64][CHUNK_BYTES];
  uint64_t full[NSTAGE];
  uint64_t empty[NSTAGE];
  uint64_t v_full;
  uint64_t s_full;
  uint64_t p_full;
  uint64_t o_full;
  float xmax[2][BT];  // row maximum / row sum of each column half
  float xsum[2][BT];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(NTHREADS, 1)
cross_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qobj, const __grid_constant__ CUtensorMap tm_qpos,
                      const __grid_constant__ CUtensorMap tm_kenc, const __grid_constant__ CUtensorMap tm_kpos,
                      const __grid_constant__ CUtensorMap tm_v, const uint32_t* __restrict__ mask_bits,
                      int words_per_row, float* __restrict__ ws_o, float* __restrict__ ws_ml, int Q, int N, int nqt,
                      float scale_log2, Drop dp) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, nkv = gridDim.x;
  const int br = blockIdx.y / nqt, qt = blockIdx.y - br * nqt;
  const int b = blockIdx.z;
  const int qrow0 = b * Q + qt * BT;
  const int krow0 = b * N + j * BT;
  constexpr uint32_t C_O = 128;

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.v_full, 1);
    mbar_init(&sm.s_full, 1);
    mbar_init(&sm.p_full, 256);
    mbar_init(&sm.o_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 8) {
    if (elect_one()) {
      for (int c = 0; c < NCHUNK; ++c) {
        const int s = c % NSTAGE;
        mbar_wait(&sm.empty[s], ((c / NSTAGE) & 1) ^ 1, 31);
        mbar_arrive_expect_tx(&sm.full[s], 2 * CHUNK_BYTES);
        if (c < 4) {
          tma_load_2d(sm.a[s], &tm_qobj, &sm.full[s], br * 256 + c * 64, qrow0);
          tma_load_2d(sm.b[s], &tm_kenc, &sm.full[s], c * 64, krow0);
        } else {
          tma_load_2d(sm.a[s], &tm_qpos, &sm.full[s], (c - 4) * 64, qrow0);
          tma_load_2d(sm.b[s], &tm_kpos, &sm.full[s], (c - 4) * 64, krow0);
        }
      }
      mbar_arrive_expect_tx(&sm.v_full, (DV / 64) * CHUNK_BYTES);
      for (int c = 0; c < DV / 64; ++c) tma_load_2d(sm.v[c], &tm_v, &sm.v_full, c * 64, krow0);
    }
    __syncwarp();
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t id_qk = umma_idesc_bf16(BT, BT, false, false);
      constexpr uint32_t id_pv = umma_idesc_bf16(BT, 64, false, true);
      for (int c = 0; c < NCHUNK; ++c) {
        const int s = c % NSTAGE;
        mbar_wait(&sm.full[s], (c / NSTAGE) & 1, 32);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_ss(tmem, umma_smem_desc(smem_u32(sm.a[s]) + ks * 32, 16, 1024, SWZ_128B),
                  umma_smem_desc(smem_u32(sm.b[s]) + ks * 32, 16, 1024, SWZ_128B), id_qk, (c > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&sm.empty[s]);
      }
      tc_commit(&sm.s_full);
      mbar_wait(&sm.p_full, 0, 33);
      mbar_wait(&sm.v_full, 0, 34);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DV / 64; ++c) {
#pragma unroll
        for (int ks = 0; ks < BT / 16; ++ks) {
          umma_ts(tmem + C_O + c * 64, tmem + ks * 8, umma_smem_desc(smem_u32(sm.v[c]) + ks * 2048, 16, 1024, SWZ_128B),
                  id_pv, ks > 0);
        }
      }
      tc_commit(&sm.o_full);
    }
    __syncwarp();
  } else {
    const int wq = warp & 3, ch = warp >> 2;  // TMEM lane quadrant, half of the 128 key columns
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const int r_in_tile = wq * 32 + lane;
    const int q = qt * BT + r_in_tile;
    const uint32_t drop_seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
    const uint32_t drop_row = static_cast<uint32_t>((b * 2 + br) * Q + q);
    mbar_wait(&sm.s_full, 0, 35);
    tc_fence_after();
    uint32_t sr[2][32];  // keys [64 ch, 64 ch + 64) of the tile
#pragma unroll
    for (int c = 0; c < 2; ++c) tmem_ld_x32(tmem + lane_addr + (2 * ch + c) * 32, sr[c]);
    tc_wait_ld();
    const uint2 mw = *reinterpret_cast<const uint2*>(mask_bits + static_cast<size_t>(b) * words_per_row + j * 4 + 2 * ch);
    const uint32_t mwa[2] = {mw.x, mw.y};
    if ((mw.x | mw.y) != 0u) {
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if ((mwa[c] >> i) & 1u) sr[c][i] = 0xff800000u;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sr[c][i]));
    sm.xmax[ch][r_in_tile] = mx;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float m = fmaxf(mx, sm.xmax[1 - ch][r_in_tile]) * scale_log2;
    const float m_use = (m == -INFINITY) ? 0.f : m;
    float l = 0.f;
    {
      uint32_t pk[32];  // packed P columns [32 ch, 32 ch + 32) = keys [64 ch, 64 ch + 64)
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int e = 2 * i;
        const float p0 = ex2_approx(fmaf(__uint_as_float(sr[e >> 5][e & 31]), scale_log2, -m_use));
        const float p1 = ex2_approx(fmaf(__uint_as_float(sr[(e + 1) >> 5][(e + 1) & 31]), scale_log2, -m_use));
        l += p0 + p1;
        pk[i] = pack_bf16x2(p0, p1);
        if (dp.thr16) {  // ClsRegBranch's SelfAttention drops P (always, self_attention.py:40): row = (b, branch, query)
          const uint32_t bits = drop_bits(drop_seed, dp.site, drop_row, j * (BT / 2) + ch * 32 + i);
          pk[i] = pack_bf16x2(((bits & 0xFFFFu) >= dp.thr16) ? p0 : 0.f, ((bits >> 16) >= dp.thr16) ? p1 : 0.f);
        }
      }
      tmem_st_x32(tmem + lane_addr + ch * 32, pk);
    }
    sm.xsum[ch][r_in_tile] = l;
    tc_wait_st();
    tc_fence_before();
    mbar_arrive(&sm.p_full);
    mbar_wait(&sm.o_full, 0, 36);
    tc_fence_after();
    const size_t slot = ((static_cast<size_t>(b) * 2 + br) * nqt + qt) * nkv + j;
    float* od = ws_o + (slot * BT + r_in_tile) * DV;
#pragma unroll
    for (int c = ch * (DV / 64); c < (ch + 1) * (DV / 64); ++c) {  // this thread's half of the row's 256 output columns
      uint32_t r[32];
      tmem_ld_x32(tmem + lane_addr + C_O + c * 32, r);
      tc_wait_ld();
      if (q < Q) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<uint4*>(od + c * 32)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
      }
    }
    // (every thread passed the mbarrier wait on o_full after all 256 had arrived on p_full, i.e. after both halves of
    // xsum were written)
    if (ch == 0 && q < Q)
      reinterpret_cast<float2*>(ws_ml)[slot * BT + r_in_tile] = make_float2(m, l + sm.xsum[1][r_in_tile]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// one warp per (b, branch, query): merge the nkv partials
__global__ void cross_attn_combine_kernel(const float* __restrict__ ws_o, const float* __restrict__ ws_ml,
                                          __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int B, int Q,
                                          int nqt, int nkv, float out_scale) {
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (idx >= B * 2 * Q) return;
  const int q = idx % Q, br = (idx / Q) & 1, b = idx / (2 * Q);
  const int qt = q / BT, r = q - qt * BT;
  const size_t slot0 = ((static_cast<size_t>(b) * 2 + br) * nqt + qt) * nkv;
  float M = -INFINITY;
  for (int j = 0; j < nkv; ++j) M = fmaxf(M, ws_ml[((slot0 + j) * BT + r) * 2]);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float L = 0.f;
  for (int j = 0; j < nkv; ++j) {
    const float2 ml = reinterpret_cast<const float2*>(ws_ml)[(slot0 + j) * BT + r];
    const float w = ex2_approx(ml.x - M);
    L = fmaf(w, ml.y, L);
    const float4* src = reinterpret_cast<const float4*>(ws_o + ((slot0 + j) * BT + r) * DV + lane * 8);
    const float4 x0 = src[0], x1 = src[1];
    acc[0] = fmaf(w, x0.x, acc[0]); acc[1] = fmaf(w, x0.y, acc[1]); acc[2] = fmaf(w, x0.z, acc[2]);
    acc[3] = fmaf(w, x0.w, acc[3]); acc[4] = fmaf(w, x1.x, acc[4]); acc[5] = fmaf(w, x1.y, acc[5]);
    acc[6] = fmaf(w, x1.z, acc[6]); acc[7] = fmaf(w, x1.w, acc[7]);
  }
  const float inv = out_scale / L;  // out_scale = 1/(1-p) of the attention-probability dropout
  const uint4 o = make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                             pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
  *reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * Q + q) * 512 + br * 256 + lane * 8) = o;
  if (lse && lane == 0) lse[(static_cast<size_t>(b) * 2 + br) * Q + q] = M + lg2_approx(L);
}

}  // namespace
}  // namespace destr

extern "C" int64_t destr_split_cross_attn_ws_floats(int B, int Q, int N) {
  const int64_t nqt = (Q + 127) / 128, nkv = (N + 127) / 128;
  return static_cast<int64_t>(B) * 2 * nqt * nkv * 128 * (256 + 2);
}

extern "C" int destr_split_cross_attn_fwd(const void* q_obj, const void* q_pos, const void* k_enc, const void* k_pos,
                                          const void* v, int ld_kenc, int ld_kpos, int ld_v,
                                          const uint32_t* mask_bits, int words_per_row, void* out, float* lse,
                                          float* ws_partial, int B, int Q, int N, float scale,
                                          const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site,
                                          void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(q_obj && q_pos && k_enc && k_pos && v && mask_bits && out && ws_partial, "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && N > 0, "shape");
  const int nqt = ceil_div(Q, BT), nkv = ceil_div(N, BT);
  DESTR_CHECK_ARG(words_per_row >= nkv * 4, "words_per_row");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t qrows = static_cast<uint64_t>(B) * Q, krows = static_cast<uint64_t>(B) * N;
  CUtensorMap tqo, tqp, tke, tkp, tv;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tqo, q_obj, qrows, 512, 512, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tqp, q_pos, qrows, 256, 256, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tke, k_enc, krows, 256, ld_kenc, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tkp, k_pos, krows, 256, ld_kpos, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, v, krows, 256, ld_v, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const size_t smem = sizeof(Smem) + 1024;
  DESTR_SMEM_OPTIN(cross_attn_fwd_kernel, smem);
  float* ws_o = ws_partial;
  float* ws_ml = ws_partial + static_cast<size_t>(B) * 2 * nqt * nkv * BT * DV;
  dim3 grid(nkv, 2 * nqt, B);
  cross_attn_fwd_kernel<<<grid, NTHREADS, smem, st>>>(tqo, tqp, tke, tkp, tv, mask_bits, words_per_row, ws_o, ws_ml,
                                                      Q, N, nqt, scale * 1.4426950408889634f,
                                                      Drop{drop_seed, drop_thr16, drop_site});
  DESTR_LAUNCH_CHECK();
  const int total = B * 2 * Q;
  cross_attn_combine_kernel<<<ceil_div(total, 8), 256, 0, st>>>(ws_o, ws_ml, static_cast<__nv_bfloat16*>(out), lse, B,
                                                               Q, nqt, nkv, drop_scale(drop_thr16));
  DESTR_LAUNCH_CHECK();
  return 0;
}
