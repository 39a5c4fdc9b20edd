// Decoder self-attention + pair self-attention forward, one launch, on tcgen05 / TMEM / TMA.
//
// Reference arithmetic: SelfAttention.forward (src/model/attention/self_attention.py:26-45, 8 heads x
// 64, scale 1/sqrt(64)) and PairSelfAttention.forward (src/model/attention/pair_self_attention.py:
// 91-99: A2 = Ql.Kl^T + Qr.Kr^T, P = softmax(A2) / sqrt(2*64), O = P.[Vl|Vr]).  With the left/right
// gathers done up front (destr_dec_qkv_prep writes qcat = [q[L] | q[R]] etc.), pair attention is a
// plain attention with d_head = 128 whose softmax is divided by sqrt(128) AFTER normalisation.
//
// grid = (ceil(Q/128), 16, B): blockIdx.y < 8 -> self-attention head (D = 64), else pair head (D = 128).
// The whole key range (Q <= 384) fits on chip, so there is no online softmax: S for all key tiles sits in
// TMEM (up to 3 x 128 fp32 columns), softmax threads take the row max, write P (bf16) over S, and
// O = sum_j P_j.V_j accumulates in TMEM ([384, 384+D)).  V_j is TMA-loaded into K_j's smem slot once
// Q.K_j^T has retired.  Every mbarrier is used exactly once (no phases).
// warps 0-3 softmax (thread <-> query row <-> TMEM lane), warp 4 TMA, warp 5 MMA.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {
namespace {

constexpr int BT = 128;
constexpr int MAX_TILES = 3;
constexpr int NTHREADS = 192;
constexpr uint32_t CHUNK_BYTES = BT * 128;  // [128 rows][64 bf16], SW128 : 16 KB

struct __align__(1024) Smem {
  uint8_t q[2][CHUNK_BYTES];              // D/64 chunks
  uint8_t kv[MAX_TILES][2][CHUNK_BYTES];  // K_j, later V_j
  uint64_t q_full;
  uint64_t k_full[MAX_TILES];
  uint64_t k_free[MAX_TILES];
  uint64_t v_full[MAX_TILES];
  uint64_t s_full;
  uint64_t p_full;
  uint64_t o_full;
  uint32_t tmem_base;
};

struct Params {
  __nv_bfloat16* o1;
  __nv_bfloat16* o2;
  float* lse1;
  float* lse2;
  int Q;
  Drop dp;  // attention-probability dropout of the SELF-attention heads (self_attention.py:40); the pair heads have none
};

template <int D>
__device__ __forceinline__ void body(Smem& sm, const CUtensorMap* tm_q, const CUtensorMap* tm_k,
                                     const CUtensorMap* tm_v, __nv_bfloat16* __restrict__ out,
                                     float* __restrict__ lse, int Q, int h, int b, int mt, float scale_log2,
                                     float out_scale, uint32_t tmem, Drop dp) {
  constexpr int NCH = D / 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkv = (Q + BT - 1) / BT;
  const int row_base = b * Q;            // token-major output rows
  const int hrow = (b * 8 + h) * Q;      // head-major operand rows: [B, 8, Q, D]
  constexpr uint32_t C_O = 384;

  if (warp == 4) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&sm.q_full, NCH * CHUNK_BYTES);
      for (int c = 0; c < NCH; ++c) tma_load_2d(sm.q[c], tm_q, &sm.q_full, c * 64, hrow + mt * BT);
      for (int j = 0; j < nkv; ++j) {
        mbar_arrive_expect_tx(&sm.k_full[j], NCH * CHUNK_BYTES);
        for (int c = 0; c < NCH; ++c) tma_load_2d(sm.kv[j][c], tm_k, &sm.k_full[j], c * 64, hrow + j * BT);
      }
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&sm.k_free[j], 0, 21);
        mbar_arrive_expect_tx(&sm.v_full[j], NCH * CHUNK_BYTES);
        for (int c = 0; c < NCH; ++c) tma_load_2d(sm.kv[j][c], tm_v, &sm.v_full[j], c * 64, hrow + j * BT);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t id_qk = umma_idesc_bf16(BT, BT, false, false);
      constexpr uint32_t id_pv = umma_idesc_bf16(BT, 64, false, true);
      mbar_wait(&sm.q_full, 0, 22);
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&sm.k_full[j], 0, 23);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          umma_ss(tmem + j * BT, umma_smem_desc(smem_u32(sm.q[ks >> 2]) + (ks & 3) * 32, 16, 1024, SWZ_128B),
                  umma_smem_desc(smem_u32(sm.kv[j][ks >> 2]) + (ks & 3) * 32, 16, 1024, SWZ_128B), id_qk, ks > 0);
        }
        tc_commit(&sm.k_free[j]);
      }
      tc_commit(&sm.s_full);
      mbar_wait(&sm.p_full, 0, 24);
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&sm.v_full[j], 0, 25);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
          for (int ks = 0; ks < BT / 16; ++ks) {
            umma_ts(tmem + C_O + c * 64, tmem + j * BT + ks * 8,
                    umma_smem_desc(smem_u32(sm.kv[j][c]) + ks * 2048, 16, 1024, SWZ_128B), id_pv,
                    (j > 0 || ks > 0) ? 1u : 0u);
          }
        }
      }
      tc_commit(&sm.o_full);
    }
    __syncwarp();
  } else {
    const int wq = warp;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const int qrow = mt * BT + wq * 32 + lane;
    const uint32_t drop_seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
    mbar_wait(&sm.s_full, 0, 26);
    tc_fence_after();
    // pass 1: row max over all keys < Q
    float mx = -INFINITY;
    for (int j = 0; j < nkv; ++j) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tmem + lane_addr + j * BT + c * 32, r);
        tc_wait_ld();
        const int key0 = j * BT + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (key0 + i < Q) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
    }
    const float m = mx * scale_log2;
    // pass 2: p = exp2(c*s - m), row sum, P (bf16) over the S columns already consumed
    float l = 0.f;
    for (int j = 0; j < nkv; ++j) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tmem + lane_addr + j * BT + c * 32, r);
        tc_wait_ld();
        const int key0 = j * BT + c * 32;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), scale_log2, -m));
          float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -m));
          p0 = (key0 + 2 * i < Q) ? p0 : 0.f;
          p1 = (key0 + 2 * i + 1 < Q) ? p1 : 0.f;
          l += p0 + p1;
          if (dp.thr16) {  // mask row = (b, h, query), column = key; the denominator l is not dropped
            const uint32_t bits = drop_bits(drop_seed, dp.site, static_cast<uint32_t>(hrow + qrow), (key0 >> 1) + i);
            if ((bits & 0xFFFFu) < dp.thr16) p0 = 0.f;
            if ((bits >> 16) < dp.thr16) p1 = 0.f;
          }
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_x16(tmem + lane_addr + j * BT + c * 16, pk);
      }
    }
    tc_wait_st();
    tc_fence_before();
    mbar_arrive(&sm.p_full);
    mbar_wait(&sm.o_full, 0, 27);
    tc_fence_after();
    const float inv = out_scale * drop_scale(dp.thr16) / l;
    const bool valid = qrow < Q;
    __nv_bfloat16* dst = out + (static_cast<size_t>(row_base + qrow) * 8 + h) * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem + lane_addr + C_O + c * 32, r);
      tc_wait_ld();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            w[e] = pack_bf16x2(__uint_as_float(r[i * 8 + 2 * e]) * inv, __uint_as_float(r[i * 8 + 2 * e + 1]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    if (valid && lse) lse[(static_cast<size_t>(b) * 8 + h) * Q + qrow] = m + lg2_approx(l);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
dec_attn_fwd_kernel(const __grid_constant__ CUtensorMap tq1, const __grid_constant__ CUtensorMap tk1,
                    const __grid_constant__ CUtensorMap tv1, const __grid_constant__ CUtensorMap tq2,
                    const __grid_constant__ CUtensorMap tk2, const __grid_constant__ CUtensorMap tv2, Params p) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    mbar_init(&sm.q_full, 1);
    for (int j = 0; j < MAX_TILES; ++j) {
      mbar_init(&sm.k_full[j], 1);
      mbar_init(&sm.k_free[j], 1);
      mbar_init(&sm.v_full[j], 1);
    }
    mbar_init(&sm.s_full, 1);
    mbar_init(&sm.p_full, 128);
    mbar_init(&sm.o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const int mt = blockIdx.x, hy = blockIdx.y, b = blockIdx.z;
  const float log2e = 1.4426950408889634f;
  if (hy < 8)
    body<64>(sm, &tq1, &tk1, &tv1, p.o1, p.lse1, p.Q, hy, b, mt, log2e * 0.125f, 1.0f, tmem, p.dp);
  else
    body<128>(sm, &tq2, &tk2, &tv2, p.o2, p.lse2, p.Q, hy - 8, b, mt, log2e, 0.08838834764831845f, tmem,
              Drop{nullptr, 0u, 0u});
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// prep: q = q_obj + [qp|qp], k = k_obj + [kp|kp] (decoder_block.py:167-177) and the left/right
// gathers of pair attention (pair_self_attention.py:47-89):  xcat[i, h*128 + 0..63] = x[L_i, h*64..],
// xcat[i, h*128 + 64..127] = x[R_i, h*64..].   One block per (b, i); 64 threads x 8 channels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 add_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  const uint32_t* pa = &a.x;
  const uint32_t* pb = &b.x;
  uint32_t* pr = &r.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lo = __uint_as_float(pa[i] << 16) + __uint_as_float(pb[i] << 16);
    const float hi = __uint_as_float(pa[i] & 0xffff0000u) + __uint_as_float(pb[i] & 0xffff0000u);
    pr[i] = pack_bf16x2(lo, hi);
  }
  return r;
}

__global__ void dec_qkv_prep_kernel(const __nv_bfloat16* __restrict__ qkv_obj,  // [B*Q, 1536] = q|k|v
                                    const __nv_bfloat16* __restrict__ qk_pos,   // [B*Q, 512] = qp|kp (pitch ld_pos)
                                    const int32_t* __restrict__ pairs,
                                    __nv_bfloat16* __restrict__ qkv,  // [3][B, 8, Q, 64]   head-major
                                    __nv_bfloat16* __restrict__ cat,  // [3][B, 8, Q, 128]  head-major [left|right]
                                    int Q, int rows, int ld_pos) {
  const int row = blockIdx.x;
  const int b = row / Q, i = row - b * Q;
  const int t = threadIdx.x;  // 0..63 -> channel t*8 of 512
  const int L = b * Q + pairs[2 * row], R = b * Q + pairs[2 * row + 1];
  auto qk_at = [&](int r, int which) {  // which: 0 q, 1 k
    const uint4 o = *reinterpret_cast<const uint4*>(qkv_obj + static_cast<size_t>(r) * 1536 + which * 512 + t * 8);
    const uint4 p = *reinterpret_cast<const uint4*>(qk_pos + static_cast<size_t>(r) * ld_pos + which * 256 + (t & 31) * 8);
    return add_bf16x8(o, p);
  };
  auto v_at = [&](int r) {
    return *reinterpret_cast<const uint4*>(qkv_obj + static_cast<size_t>(r) * 1536 + 1024 + t * 8);
  };
  const int hh = t >> 3, within = (t & 7) * 8;  // head, channel within head
  const size_t hm_row = static_cast<size_t>(b * 8 + hh) * Q + i;
  const size_t self_stride = static_cast<size_t>(rows) * 512, cat_stride = static_cast<size_t>(rows) * 1024;
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    const uint4 self = which < 2 ? qk_at(row, which) : v_at(row);
    *reinterpret_cast<uint4*>(qkv + which * self_stride + hm_row * 64 + within) = self;
    const uint4 left = which < 2 ? qk_at(L, which) : v_at(L);
    const uint4 right = which < 2 ? qk_at(R, which) : v_at(R);
    __nv_bfloat16* dst = cat + which * cat_stride + hm_row * 128 + within;
    *reinterpret_cast<uint4*>(dst) = left;
    *reinterpret_cast<uint4*>(dst + 64) = right;
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_dec_qkv_prep(const void* qkv_obj, const void* qk_pos, int ld_pos, const int32_t* pairs,
                                  void* qkv, void* cat, int B, int Q, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(qkv_obj && qk_pos && pairs && qkv && cat && B > 0 && Q > 0, "null pointer / shape");
  DESTR_CHECK_ARG(ld_pos >= 512 && ld_pos % 8 == 0, "ld_pos");
  dec_qkv_prep_kernel<<<B * Q, 64, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv_obj), static_cast<const __nv_bfloat16*>(qk_pos), pairs,
      static_cast<__nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(cat), Q, B * Q, ld_pos);
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_dec_self_pair_attn_fwd(const void* qkv, const void* cat, void* o1, void* o2, float* lse1,
                                            float* lse2, int B, int Q, const uint32_t* drop_seed, uint32_t drop_thr16,
                                            uint32_t drop_site, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(qkv && cat && o1 && o2, "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && Q <= BT * MAX_TILES, "Q must be <= 384");
  const uint64_t rows = static_cast<uint64_t>(B) * Q;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* c = static_cast<const __nv_bfloat16*>(cat);
  CUtensorMap t[6];
  int rc;
  for (int w = 0; w < 3; ++w) {  // head-major operands: 2-D [B*8*Q, D]
    if ((rc = make_tmap_bf16_2d(&t[w], x + w * rows * 512, rows * 8, 64, 64, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
    if ((rc = make_tmap_bf16_2d(&t[3 + w], c + w * rows * 1024, rows * 8, 128, 128, BT, 64,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  const size_t smem = sizeof(Smem) + 1024;
  DESTR_SMEM_OPTIN(dec_attn_fwd_kernel, smem);
  Params p{static_cast<__nv_bfloat16*>(o1), static_cast<__nv_bfloat16*>(o2), lse1, lse2, Q,
           Drop{drop_seed, drop_thr16, drop_site}};
  dim3 grid(ceil_div(Q, BT), 16, B);
  dec_attn_fwd_kernel<<<grid, NTHREADS, smem, static_cast<cudaStream_t>(stream)>>>(t[0], t[1], t[2], t[3], t[4],
                                                                                   t[5], p);
  DESTR_LAUNCH_CHECK();
  return 0;
}
