// Encoder multi-head self-attention backward (d_head = 32) on tcgen05 / TMEM / TMA.
//
// Backward of the fused attention in enc_attn_fwd.cu (reference arithmetic: torch
// F.multi_head_attention_forward as called at src/model/blocks/encoder_block.py:97-103).
// Scores are recomputed from Q,K and the saved log-sum-exp; nothing N x N is ever stored.
//
// One CTA = one 128-key tile j of one (batch, head); it loops over all 128-query tiles i.  Everything
// is computed TRANSPOSED (rows = keys) so that thread t <-> key row t <-> TMEM lane t:
//   S^T  = K_j . Q_i^T            (SS MMA, K-major x K-major)       -> TMEM ST   [0,128)
//   dP^T = V_j . dO_i^T           (SS MMA)                           -> TMEM DPT  [128,256)
//   P^T  = exp2(c*S^T - lse_q)    dS^T = P^T o (dP^T - delta_q)      (8 compute warps)
//   dV_j += P^T . dO_i            (TS MMA, A = P^T bf16 in TMEM [256,320), B = dO_i MN-major)
//   dK_j += dS^T . Q_i            (SS MMA, A = dS^T in smem K-major SW128, B = Q_i MN-major)
//   dQ_i  = dS . K_j              (SS MMA, A = the same smem dS^T read MN-major, B = K_j MN-major)
// dV_j / dK_j accumulate in TMEM ([320,352) / [352,384)) across the whole loop; dQ_i tiles
// (double buffered at [384,416) / [416,448)) are drained with red.global.add.v4.f32 into an fp32
// accumulator that a small kernel converts to bf16 afterwards.
// Warps: 0-7 compute (warp w: TMEM lanes 32*(w%4).., query columns 64*(w/4)..), 8 TMA, 9 MMA.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {
extern int g_knobs[16];
namespace {

constexpr int DH = 32;
constexpr int BT = 128;  // tile edge (keys and queries)
constexpr int QSTAGES = 2;
constexpr int NTHREADS = 320;
constexpr uint32_t TILE_BYTES = BT * DH * 2;  // 8192
constexpr uint32_t DS_BLOCK_BYTES = BT * 128;  // 16384: [128 key rows][64 queries] bf16, SW128

struct __align__(1024) Smem {
  uint8_t k[TILE_BYTES];
  uint8_t v[TILE_BYTES];
  uint8_t q[QSTAGES][TILE_BYTES];
  uint8_t d_o[QSTAGES][TILE_BYTES];
  uint8_t ds[2][DS_BLOCK_BYTES];
  float lse_s[2][BT];
  float delta_s[2][BT];
  uint64_t kv_full;
  uint64_t q_full[QSTAGES];
  uint64_t q_empty[QSTAGES];
  uint64_t sdp_full;
  uint64_t pds_full;
  uint64_t dq_full[2];
  uint64_t dkv_full;
  uint32_t tmem_base;
};

struct Knobs {
  uint32_t mn64_lbo, mn64_sbo, kmaj_lbo, a_mn_lbo, a_mn_sbo, a_mn_kstep;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
enc_attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                    const uint32_t* __restrict__ mask_bits, int words_per_row, const float* __restrict__ lse,
                    const float* __restrict__ delta, float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dk,
                    __nv_bfloat16* __restrict__ dv, int ld_dk, int ld_dv, int N, int heads, float scale,
                    float scale_log2, Knobs kn) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nq = (N + BT - 1) / BT;
  const int row_base = b * N;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(&sm.kv_full, 1);
    for (int s = 0; s < QSTAGES; ++s) {
      mbar_init(&sm.q_full[s], 1);
      mbar_init(&sm.q_empty[s], 1);
    }
    mbar_init(&sm.sdp_full, 1);
    mbar_init(&sm.pds_full, 256);
    mbar_init(&sm.dq_full[0], 1);
    mbar_init(&sm.dq_full[1], 1);
    mbar_init(&sm.dkv_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  constexpr uint32_t C_ST = 0, C_DPT = 128, C_PT = 256, C_DV = 320, C_DK = 352, C_DQ = 384;

  if (warp == 8) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&sm.kv_full, 2 * TILE_BYTES);
      tma_load_2d(sm.k, &tm_k, &sm.kv_full, h * DH, row_base + j * BT);
      tma_load_2d(sm.v, &tm_v, &sm.kv_full, h * DH, row_base + j * BT);
      for (int i = 0; i < nq; ++i) {
        const int s = i % QSTAGES;
        mbar_wait(&sm.q_empty[s], ((i / QSTAGES) & 1) ^ 1, 11);
        mbar_arrive_expect_tx(&sm.q_full[s], 2 * TILE_BYTES);
        tma_load_2d(sm.q[s], &tm_q, &sm.q_full[s], h * DH, row_base + i * BT);
        tma_load_2d(sm.d_o[s], &tm_do, &sm.q_full[s], h * DH, row_base + i * BT);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t id_sT = umma_idesc_bf16(BT, BT, false, false);   // K-major x K-major, N=128
      constexpr uint32_t id_kn = umma_idesc_bf16(BT, DH, false, true);    // A K-major (TMEM/smem), B MN-major
      constexpr uint32_t id_nn = umma_idesc_bf16(BT, DH, true, true);     // A MN-major, B MN-major
      mbar_wait(&sm.kv_full, 0, 12);
      for (int i = 0; i < nq; ++i) {
        const int s = i % QSTAGES;
        mbar_wait(&sm.q_full[s], (i / QSTAGES) & 1, 13);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          umma_ss(tmem + C_ST, umma_smem_desc(smem_u32(sm.k) + ks * 32, kn.kmaj_lbo, 512, SWZ_64B),
                  umma_smem_desc(smem_u32(sm.q[s]) + ks * 32, kn.kmaj_lbo, 512, SWZ_64B), id_sT, ks > 0);
        }
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          umma_ss(tmem + C_DPT, umma_smem_desc(smem_u32(sm.v) + ks * 32, kn.kmaj_lbo, 512, SWZ_64B),
                  umma_smem_desc(smem_u32(sm.d_o[s]) + ks * 32, kn.kmaj_lbo, 512, SWZ_64B), id_sT, ks > 0);
        }
        tc_commit(&sm.sdp_full);
        mbar_wait(&sm.pds_full, i & 1, 14);
        tc_fence_after();
        // dV += P^T . dO_i
#pragma unroll
        for (int ks = 0; ks < BT / 16; ++ks) {
          umma_ts(tmem + C_DV, tmem + C_PT + ks * 8,
                  umma_smem_desc(smem_u32(sm.d_o[s]) + ks * 1024, kn.mn64_lbo, kn.mn64_sbo, SWZ_64B), id_kn,
                  (i > 0 || ks > 0) ? 1u : 0u);
        }
        // dK += dS^T . Q_i
#pragma unroll
        for (int ks = 0; ks < BT / 16; ++ks) {
          umma_ss(tmem + C_DK,
                  umma_smem_desc(smem_u32(sm.ds[ks >> 2]) + (ks & 3) * 32, kn.kmaj_lbo, 1024, SWZ_128B),
                  umma_smem_desc(smem_u32(sm.q[s]) + ks * 1024, kn.mn64_lbo, kn.mn64_sbo, SWZ_64B), id_kn,
                  (i > 0 || ks > 0) ? 1u : 0u);
        }
        // dQ_i = dS . K_j
#pragma unroll
        for (int ks = 0; ks < BT / 16; ++ks) {
          umma_ss(tmem + C_DQ + (i & 1) * 32,
                  umma_smem_desc(smem_u32(sm.ds[0]) + ks * kn.a_mn_kstep, kn.a_mn_lbo, kn.a_mn_sbo, SWZ_128B),
                  umma_smem_desc(smem_u32(sm.k) + ks * 1024, kn.mn64_lbo, kn.mn64_sbo, SWZ_64B), id_nn, ks > 0);
        }
        tc_commit(&sm.dq_full[i & 1]);
        tc_commit(&sm.q_empty[s]);
      }
      tc_commit(&sm.dkv_full);
    }
    __syncwarp();
  } else {
    // ------------------------------ compute warps ------------------------------
    const int wq = warp & 3, half = warp >> 2;
    const int tid = threadIdx.x;  // 0..255
    const int krow = wq * 32 + lane;
    const int key = j * BT + krow;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const bool key_masked =
        (mask_bits[static_cast<size_t>(b) * words_per_row + (key >> 5)] >> (key & 31)) & 1u;
    const float* lse_bh = lse + (static_cast<size_t>(b) * heads + h) * N;
    const float* delta_bh = delta + (static_cast<size_t>(b) * heads + h) * N;

    auto drain_dq = [&](int i) {
      mbar_wait(&sm.dq_full[i & 1], (i >> 1) & 1, 16);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_x16(tmem + lane_addr + C_DQ + (i & 1) * 32 + half * 16, r);
      tc_wait_ld();
      const int qrow = i * BT + krow;  // dQ tile rows are queries
      if (qrow < N) {
        float* dst = dq_acc + (static_cast<size_t>(row_base + qrow) * heads + h) * DH + half * 16;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          red_add_v4(dst + 4 * c, __uint_as_float(r[4 * c]) * scale, __uint_as_float(r[4 * c + 1]) * scale,
                     __uint_as_float(r[4 * c + 2]) * scale, __uint_as_float(r[4 * c + 3]) * scale);
      }
    };

    for (int i = 0; i < nq; ++i) {
      {
        const int qq = i * BT + (tid & 127);
        if (tid < 128)
          sm.lse_s[i & 1][tid] = (qq < N) ? lse_bh[qq] : INFINITY;
        else
          sm.delta_s[i & 1][tid - 128] = (qq < N) ? delta_bh[qq] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&sm.sdp_full, i & 1, 15);
      tc_fence_after();
      uint32_t st[2][32], dp[2][32];
      tmem_ld_x32(tmem + lane_addr + C_ST + half * 64, st[0]);
      tmem_ld_x32(tmem + lane_addr + C_ST + half * 64 + 32, st[1]);
      tmem_ld_x32(tmem + lane_addr + C_DPT + half * 64, dp[0]);
      tmem_ld_x32(tmem + lane_addr + C_DPT + half * 64 + 32, dp[1]);
      tc_wait_ld();
      const float4* lse4 = reinterpret_cast<const float4*>(&sm.lse_s[i & 1][half * 64]);
      const float4* del4 = reinterpret_cast<const float4*>(&sm.delta_s[i & 1][half * 64]);
      uint32_t pk[32];
      uint8_t* ds_row = sm.ds[half] + krow * 128;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {  // 8 queries per 16-byte chunk
        float pv[8], dsv[8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 l4 = lse4[c8 * 2 + u];
          const float4 d4 = del4[c8 * 2 + u];
          const float la[4] = {l4.x, l4.y, l4.z, l4.w};
          const float da[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c8 * 8 + u * 4 + e;
            float p = ex2_approx(fmaf(__uint_as_float(st[c >> 5][c & 31]), scale_log2, -la[e]));
            p = key_masked ? 0.f : p;
            pv[u * 4 + e] = p;
            dsv[u * 4 + e] = p * (__uint_as_float(dp[c >> 5][c & 31]) - da[e]);
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) pk[c8 * 4 + e] = pack_bf16x2(pv[2 * e], pv[2 * e + 1]);
        const uint4 dsq = make_uint4(pack_bf16x2(dsv[0], dsv[1]), pack_bf16x2(dsv[2], dsv[3]),
                                     pack_bf16x2(dsv[4], dsv[5]), pack_bf16x2(dsv[6], dsv[7]));
        *reinterpret_cast<uint4*>(ds_row + ((c8 ^ (krow & 7)) << 4)) = dsq;
      }
      tmem_st_x32(tmem + lane_addr + C_PT + half * 32, pk);
      tc_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&sm.pds_full);
      if (i > 0) drain_dq(i - 1);
    }
    drain_dq(nq - 1);
    mbar_wait(&sm.dkv_full, 0, 17);
    tc_fence_after();
    uint32_t rv[16], rk[16];
    tmem_ld_x16(tmem + lane_addr + C_DV + half * 16, rv);
    tmem_ld_x16(tmem + lane_addr + C_DK + half * 16, rk);
    tc_wait_ld();
    if (key < N) {
      uint32_t ov[8], ok[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        ov[c] = pack_bf16x2(__uint_as_float(rv[2 * c]), __uint_as_float(rv[2 * c + 1]));
        ok[c] = pack_bf16x2(__uint_as_float(rk[2 * c]) * scale, __uint_as_float(rk[2 * c + 1]) * scale);
      }
      uint4* pv_ = reinterpret_cast<uint4*>(dv + static_cast<size_t>(row_base + key) * ld_dv + h * DH + half * 16);
      uint4* pk_ = reinterpret_cast<uint4*>(dk + static_cast<size_t>(row_base + key) * ld_dk + h * DH + half * 16);
      pv_[0] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
      pv_[1] = make_uint4(ov[4], ov[5], ov[6], ov[7]);
      pk_[0] = make_uint4(ok[0], ok[1], ok[2], ok[3]);
      pk_[1] = make_uint4(ok[4], ok[5], ok[6], ok[7]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// delta[b,h,n] = sum_d dO[n,h,d] * O[n,h,d]   (one warp per token row; 4 lanes per head)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                  float* __restrict__ delta, int B, int N, int heads) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B * N) return;
  const int cols = heads * DH;
  float s = 0.f;
  if (lane * 8 < cols) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + static_cast<size_t>(row) * cols + lane * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(d_o + static_cast<size_t>(row) * cols + lane * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(gw[i] << 16), s);
      s = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(gw[i] & 0xffff0000u), s);
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  const int hh = lane >> 2;
  if ((lane & 3) == 0 && hh < heads) {
    const int bb = row / N, n = row - bb * N;
    delta[(static_cast<size_t>(bb) * heads + hh) * N + n] = s;
  }
}

// dq (bf16, row pitch ld) = dq_acc (fp32, dense [rows, cols])
__global__ void cvt_f32_bf16_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows,
                                         int cols, int ld) {
  const int64_t n4 = static_cast<int64_t>(rows) * cols / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 f = reinterpret_cast<const float4*>(src)[i];
    const int64_t e = i * 4;
    const int r = static_cast<int>(e / cols), c = static_cast<int>(e - static_cast<int64_t>(r) * cols);
    uint2 o2 = make_uint2(pack_bf16x2(f.x, f.y), pack_bf16x2(f.z, f.w));
    *reinterpret_cast<uint2*>(dst + static_cast<size_t>(r) * ld + c) = o2;
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_enc_attn_bwd(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v,
                                  const uint32_t* mask_bits, int words_per_row, const void* out, const void* dout,
                                  const float* lse, float* delta, float* dq_acc, void* dq, void* dk, void* dv,
                                  int ld_dq, int ld_dk, int ld_dv, int B, int N, int heads, float scale,
                                  void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(q && k && v && mask_bits && out && dout && lse && delta && dq_acc && dq && dk && dv, "null pointer");
  DESTR_CHECK_ARG(B > 0 && N > 0 && heads > 0 && heads * DH <= 256, "shape");
  DESTR_CHECK_ARG(words_per_row >= ceil_div(N, BT) * 4, "words_per_row");
  DESTR_CHECK_ARG(ld_dq % 8 == 0 && ld_dk % 8 == 0 && ld_dv % 8 == 0, "gradient row pitch must be a multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t rows = static_cast<uint64_t>(B) * N;
  const int cols = heads * DH;
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, q, rows, cols, ld_q, BT, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, k, rows, cols, ld_k, BT, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, v, rows, cols, ld_v, BT, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, dout, rows, cols, cols, BT, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  const size_t smem = sizeof(Smem) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    DESTR_CUDA(cudaFuncSetAttribute(enc_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  attn_delta_kernel<<<ceil_div((int)rows, 8), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out),
                                                            static_cast<const __nv_bfloat16*>(dout), delta, B, N,
                                                            heads);
  DESTR_LAUNCH_CHECK();
  DESTR_CUDA(cudaMemsetAsync(dq_acc, 0, rows * cols * sizeof(float), st));
  Knobs kn{(uint32_t)g_knobs[0], (uint32_t)g_knobs[1], (uint32_t)g_knobs[2],
           (uint32_t)g_knobs[6], (uint32_t)g_knobs[7], (uint32_t)g_knobs[8]};
  dim3 grid(ceil_div(N, BT), heads, B);
  enc_attn_bwd_kernel<<<grid, NTHREADS, smem, st>>>(tq, tk, tv, tdo, mask_bits, words_per_row, lse, delta, dq_acc,
                                                    static_cast<__nv_bfloat16*>(dk), static_cast<__nv_bfloat16*>(dv),
                                                    ld_dk, ld_dv, N, heads, scale, scale * 1.4426950408889634f, kn);
  DESTR_LAUNCH_CHECK();
  const int64_t n4 = static_cast<int64_t>(rows) * cols / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cvt_f32_bf16_rows_kernel<<<blocks, 256, 0, st>>>(dq_acc, static_cast<__nv_bfloat16*>(dq), (int)rows, cols, ld_dq);
  DESTR_LAUNCH_CHECK();
  return 0;
}
