// Decoder self-attention + pair self-attention backward, stage 1 (tcgen05 / TMEM / TMA), one launch:
// recompute S = q.k^T and dP = dO.v^T on the tensor cores, apply the softmax backward and emit P and dS
// (bf16, head-major [B,8,Q,Qp]).  The contractions dV = P^T dO, dQ = dS K, dK = dS^T Q are then plain
// batched GEMMs over contiguous head-major operands; the scatter back through the pair gathers is
// destr_dec_qkv_prep_bwd.  (Reference forward: self_attention.py:26-45, pair_self_attention.py:91-99.)
//
//   self (D = 64) :  P = 2^(c*S - lse),  c = log2e/8        dS = P (dP - delta) / 8
//   pair (D = 128):  A = 2^(c*S - lse),  c = log2e,  P = A r,   dS = A (r dP - delta),  r = 1/sqrt(128)
// with delta = rowsum(dO o O) (from destr_dual_ln_mix_bwd).  grid = (ceil(Q/128), 16, B) as in the forward.
// Per CTA: S for all key tiles in TMEM [0,384); dP_j in one 128-column buffer [384,512) that the math warps
// hand back to the MMA warp per key tile; dO reuses Q's smem slot, V_j reuses K_j's.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {
namespace {

constexpr int BT = 128;
constexpr int MAX_TILES = 3;
constexpr int NTHREADS = 192;
constexpr uint32_t CHUNK_BYTES = BT * 128;

struct __align__(1024) Smem {
  uint8_t q[2][CHUNK_BYTES];              // Q tile, then dO tile
  uint8_t kv[MAX_TILES][2][CHUNK_BYTES];  // K_j, then V_j
  uint64_t q_full, s_full, do_full;
  uint64_t k_full[MAX_TILES], k_free[MAX_TILES], v_full[MAX_TILES], dp_full[MAX_TILES], dp_free[MAX_TILES];
  uint32_t tmem_base;
};

struct Params {
  const float* lse1; const float* lse2; const float* delta1; const float* delta2;
  __nv_bfloat16* P1; __nv_bfloat16* dS1; __nv_bfloat16* P2; __nv_bfloat16* dS2;
  int Q, Qp;
  Drop dp;  // self-attention heads only (self_attention.py:40)
};

template <int D>
__device__ __forceinline__ void body(Smem& sm, const CUtensorMap* tm_q, const CUtensorMap* tm_k,
                                     const CUtensorMap* tm_v, const CUtensorMap* tm_do, const float* __restrict__ lse,
                                     const float* __restrict__ delta, __nv_bfloat16* __restrict__ P_out,
                                     __nv_bfloat16* __restrict__ dS_out, int Q, int Qp, int h, int b, int mt,
                                     float scale_log2, float p_scale, float dp_scale, float ds_scale, uint32_t tmem, Drop dp) {
  constexpr int NCH = D / 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkv = (Q + BT - 1) / BT;
  const int hrow = (b * 8 + h) * Q;
  constexpr uint32_t C_DP = 384;

  if (warp == 4) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&sm.q_full, NCH * CHUNK_BYTES);
      for (int c = 0; c < NCH; ++c) tma_load_2d(sm.q[c], tm_q, &sm.q_full, c * 64, hrow + mt * BT);
      for (int j = 0; j < nkv; ++j) {
        mbar_arrive_expect_tx(&sm.k_full[j], NCH * CHUNK_BYTES);
        for (int c = 0; c < NCH; ++c) tma_load_2d(sm.kv[j][c], tm_k, &sm.k_full[j], c * 64, hrow + j * BT);
      }
      mbar_wait(&sm.s_full, 0, 51);  // every Q.K^T has retired: Q's slot is free for dO
      mbar_arrive_expect_tx(&sm.do_full, NCH * CHUNK_BYTES);
      for (int c = 0; c < NCH; ++c) tma_load_2d(sm.q[c], tm_do, &sm.do_full, c * 64, hrow + mt * BT);
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&sm.k_free[j], 0, 52);
        mbar_arrive_expect_tx(&sm.v_full[j], NCH * CHUNK_BYTES);
        for (int c = 0; c < NCH; ++c) tma_load_2d(sm.kv[j][c], tm_v, &sm.v_full[j], c * 64, hrow + j * BT);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BT, BT, false, false);
      mbar_wait(&sm.q_full, 0, 53);
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&sm.k_full[j], 0, 54);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          umma_ss(tmem + j * BT, umma_smem_desc(smem_u32(sm.q[ks >> 2]) + (ks & 3) * 32, 16, 1024, SWZ_128B),
                  umma_smem_desc(smem_u32(sm.kv[j][ks >> 2]) + (ks & 3) * 32, 16, 1024, SWZ_128B), idesc, ks > 0);
        }
        tc_commit(&sm.k_free[j]);
      }
      tc_commit(&sm.s_full);
      mbar_wait(&sm.do_full, 0, 55);
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&sm.v_full[j], 0, 56);
        if (j > 0) mbar_wait(&sm.dp_free[j - 1], 0, 57);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {  // dP_j = dO . V_j^T  (both K-major, contraction over D)
          umma_ss(tmem + C_DP, umma_smem_desc(smem_u32(sm.q[ks >> 2]) + (ks & 3) * 32, 16, 1024, SWZ_128B),
                  umma_smem_desc(smem_u32(sm.kv[j][ks >> 2]) + (ks & 3) * 32, 16, 1024, SWZ_128B), idesc, ks > 0);
        }
        tc_commit(&sm.dp_full[j]);
      }
    }
    __syncwarp();
  } else {
    const int wq = warp;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const int qrow = mt * BT + wq * 32 + lane;
    const bool valid = qrow < Q;
    const uint32_t drop_seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
    const float drop_s = drop_scale(dp.thr16);  // 1 when dropout is off (thr16 = 0 keeps everything)
    const float l2 = valid ? lse[static_cast<size_t>(hrow) + qrow] : INFINITY;
    const float dl = valid ? delta[static_cast<size_t>(hrow) + qrow] : 0.f;
    __nv_bfloat16* prow = P_out + (static_cast<size_t>(hrow) + (valid ? qrow : 0)) * Qp;
    __nv_bfloat16* drow = dS_out + (static_cast<size_t>(hrow) + (valid ? qrow : 0)) * Qp;
    mbar_wait(&sm.s_full, 0, 58);
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&sm.dp_full[j], 0, 59);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t s[32], d[32];
        tmem_ld_x32(tmem + lane_addr + j * BT + c * 32, s);
        tmem_ld_x32(tmem + lane_addr + C_DP + c * 32, d);
        tc_wait_ld();
        const int key0 = j * BT + c * 32;
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pv[2], dv[2];
          uint32_t bits = 0xFFFFFFFFu;
          if (dp.thr16) bits = drop_bits(drop_seed, dp.site, static_cast<uint32_t>(hrow + qrow), (key0 >> 1) + i);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = 2 * i + e;
            float p = ex2_approx(fmaf(__uint_as_float(s[k]), scale_log2, -l2));
            p = (key0 + k < Q) ? p : 0.f;
            // forward: O = dropout(P) V  ->  dV takes the dropped P, dP passes through the same mask
            const float keep = (((bits >> (16 * e)) & 0xFFFFu) >= dp.thr16) ? drop_s : 0.f;
            pv[e] = p * p_scale * keep;
            dv[e] = p * (__uint_as_float(d[k]) * keep * dp_scale - dl) * ds_scale;
          }
          pp[i] = pack_bf16x2(pv[0], pv[1]);
          dd[i] = pack_bf16x2(dv[0], dv[1]);
        }
        if (valid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            reinterpret_cast<uint4*>(prow + key0)[i] = make_uint4(pp[4 * i], pp[4 * i + 1], pp[4 * i + 2], pp[4 * i + 3]);
            reinterpret_cast<uint4*>(drow + key0)[i] = make_uint4(dd[4 * i], dd[4 * i + 1], dd[4 * i + 2], dd[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&sm.dp_free[j]);
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
dec_attn_bwd_ds_kernel(const __grid_constant__ CUtensorMap tq1, const __grid_constant__ CUtensorMap tk1,
                       const __grid_constant__ CUtensorMap tv1, const __grid_constant__ CUtensorMap td1,
                       const __grid_constant__ CUtensorMap tq2, const __grid_constant__ CUtensorMap tk2,
                       const __grid_constant__ CUtensorMap tv2, const __grid_constant__ CUtensorMap td2, Params p) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    mbar_init(&sm.q_full, 1);
    mbar_init(&sm.s_full, 1);
    mbar_init(&sm.do_full, 1);
    for (int j = 0; j < MAX_TILES; ++j) {
      mbar_init(&sm.k_full[j], 1);
      mbar_init(&sm.k_free[j], 1);
      mbar_init(&sm.v_full[j], 1);
      mbar_init(&sm.dp_full[j], 1);
      mbar_init(&sm.dp_free[j], 128);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const int mt = blockIdx.x, hy = blockIdx.y, b = blockIdx.z;
  const float log2e = 1.4426950408889634f, r = 0.08838834764831845f;
  if (hy < 8)
    body<64>(sm, &tq1, &tk1, &tv1, &td1, p.lse1, p.delta1, p.P1, p.dS1, p.Q, p.Qp, hy, b, mt, log2e * 0.125f, 1.f, 1.f,
             0.125f, tmem, p.dp);
  else
    body<128>(sm, &tq2, &tk2, &tv2, &td2, p.lse2, p.delta2, p.P2, p.dS2, p.Q, p.Qp, hy - 8, b, mt, log2e, r, r, 1.f,
              tmem, Drop{nullptr, 0u, 0u});
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// FUSED backward for Q <= 128 (one query tile = one key tile, the training shape: 100 queries): the whole
// backward of one (image, head) in one CTA, nothing but the gradients leaves the chip.
//   S = Q K^T, dP = dO V^T                    tensor cores -> TMEM [0,128), [128,256)
//   P, dS                                     softmax backward in registers (thread = query row) -> bf16 tiles in smem
//   dQ = dS K        (A = dS   K-major,  B = K  MN-major)   -> TMEM [256, 256+D)
//   dK = dS^T Q      (A = dS^T MN-major, B = Q  MN-major)   -> TMEM [0, D)      (S is consumed)
//   dV = P^T dO      (A = P^T  MN-major, B = dO MN-major)   -> TMEM [128,128+D) (dP is consumed)
// The P / dS tiles are [128 query rows][128 key columns] as two SW128 chunks of 64 columns: read along the rows they
// are the K-major A of dQ, read along the columns (64-column chunk = one MN atom, LBO = chunk stride) the MN-major A
// of dK / dV -- one copy serves both, as in enc_attn_bwd.  Q, K, V, dO tiles are TMA-loaded once and serve as
// K-major operands of the first two products and as MN-major B operands of the last three.
// ------------------------------------------------------------------------------------------------
struct __align__(1024) SmemF {
  uint8_t q[2][CHUNK_BYTES], k[2][CHUNK_BYTES], v[2][CHUNK_BYTES], dO[2][CHUNK_BYTES];
  uint8_t p[2][CHUNK_BYTES], ds[2][CHUNK_BYTES];
  uint64_t qk_full, dov_full, s_full, p_full, g_full;
  uint32_t tmem_base;
};

struct ParamsF {
  const float* lse1; const float* lse2; const float* delta1; const float* delta2;
  __nv_bfloat16* dqkv;  // [3][B,8,Q,64]
  __nv_bfloat16* dcat;  // [3][B,8,Q,128]
  int Q; size_t rows;   // rows = B*Q
  Drop dp;
};

constexpr int NTHREADS_F = 320;  // fused kernel: 8 math warps + TMA warp + MMA warp

template <int D>
__device__ __forceinline__ void body_fused(SmemF& sm, const CUtensorMap* tm_q, const CUtensorMap* tm_k,
                                           const CUtensorMap* tm_v, const CUtensorMap* tm_do,
                                           const float* __restrict__ lse, const float* __restrict__ delta,
                                           __nv_bfloat16* __restrict__ dq_out, __nv_bfloat16* __restrict__ dk_out,
                                           __nv_bfloat16* __restrict__ dv_out, int Q, int h, int b, float scale_log2,
                                           float p_scale, float dp_scale, float ds_scale, uint32_t tmem, Drop dp) {
  constexpr int NCH = D / 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hrow = (b * 8 + h) * Q;
  constexpr uint32_t C_DP = 128, C_DQ = 256, C_DK = 0, C_DV = 128;

  if (warp == 8) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&sm.qk_full, 2 * NCH * CHUNK_BYTES);
      for (int c = 0; c < NCH; ++c) tma_load_2d(sm.q[c], tm_q, &sm.qk_full, c * 64, hrow);
      for (int c = 0; c < NCH; ++c) tma_load_2d(sm.k[c], tm_k, &sm.qk_full, c * 64, hrow);
      mbar_arrive_expect_tx(&sm.dov_full, 2 * NCH * CHUNK_BYTES);
      for (int c = 0; c < NCH; ++c) tma_load_2d(sm.dO[c], tm_do, &sm.dov_full, c * 64, hrow);
      for (int c = 0; c < NCH; ++c) tma_load_2d(sm.v[c], tm_v, &sm.dov_full, c * 64, hrow);
    }
    __syncwarp();
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t id_kk = umma_idesc_bf16(BT, BT, false, false);  // K-major x K-major, N = 128
      constexpr uint32_t id_kn = umma_idesc_bf16(BT, 64, false, true);   // A K-major, B MN-major, N = 64
      constexpr uint32_t id_nn = umma_idesc_bf16(BT, 64, true, true);    // A MN-major, B MN-major, N = 64
      constexpr uint64_t D_K = umma_desc_const(16, 1024, SWZ_128B);        // K-major (k-step 32 B inside the row)
      constexpr uint64_t D_BMN = umma_desc_const(16, 1024, SWZ_128B);      // MN-major, one 64-wide atom (k-step 2048 B)
      constexpr uint64_t D_AMN = umma_desc_const(16384, 1024, SWZ_128B);   // MN-major, two 64-wide atoms 16 KB apart
      mbar_wait(&sm.qk_full, 0, 71);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks)
        umma_ss(tmem, D_K + ((smem_u32(sm.q[ks >> 2]) + (ks & 3) * 32) >> 4),
                D_K + ((smem_u32(sm.k[ks >> 2]) + (ks & 3) * 32) >> 4), id_kk, ks > 0);
      mbar_wait(&sm.dov_full, 0, 72);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks)
        umma_ss(tmem + C_DP, D_K + ((smem_u32(sm.dO[ks >> 2]) + (ks & 3) * 32) >> 4),
                D_K + ((smem_u32(sm.v[ks >> 2]) + (ks & 3) * 32) >> 4), id_kk, ks > 0);
      tc_commit(&sm.s_full);
      mbar_wait(&sm.p_full, 0, 73);  // P and dS are in shared memory, S / dP have been read out of TMEM
      tc_fence_after();
      const uint32_t a_ds = smem_u32(sm.ds[0]), a_p = smem_u32(sm.p[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
#pragma unroll
        for (int ks = 0; ks < BT / 16; ++ks) {  // contraction over the 128 keys / queries, 16 at a time
          // dQ[:, 64c..] += dS[:, 16ks..] . K[16ks.., 64c..]
          umma_ss(tmem + C_DQ + c * 64, D_K + ((a_ds + (ks >> 2) * CHUNK_BYTES + (ks & 3) * 32) >> 4),
                  D_BMN + ((smem_u32(sm.k[c]) + ks * 2048) >> 4), id_kn, ks > 0);
          // dK[:, 64c..] += dS^T[:, 16ks..] . Q[16ks.., 64c..]      (A: rows of the tile are the contraction index)
          umma_ss(tmem + C_DK + c * 64, D_AMN + ((a_ds + ks * 2048) >> 4), D_BMN + ((smem_u32(sm.q[c]) + ks * 2048) >> 4),
                  id_nn, ks > 0);
          // dV[:, 64c..] += P^T[:, 16ks..] . dO[16ks.., 64c..]
          umma_ss(tmem + C_DV + c * 64, D_AMN + ((a_p + ks * 2048) >> 4), D_BMN + ((smem_u32(sm.dO[c]) + ks * 2048) >> 4),
                  id_nn, ks > 0);
        }
      }
      tc_commit(&sm.g_full);
    }
    __syncwarp();
  } else {
    // 8 math warps: warps w and w + 4 own the same TMEM lanes (rows) and split the columns (half the instruction
    // stream per thread: the kernel is one latency chain per CTA)
    const int wq = warp & 3, ch = warp >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const int r = wq * 32 + lane;  // query row in the softmax phase, output row (query / key) in the epilogue
    const bool valid = r < Q;
    const uint32_t drop_seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
    const float drop_s = drop_scale(dp.thr16);
    const float l2 = valid ? lse[static_cast<size_t>(hrow) + r] : INFINITY;
    const float dl = valid ? delta[static_cast<size_t>(hrow) + r] : 0.f;
    const uint32_t a_p = smem_u32(sm.p[0]), a_ds = smem_u32(sm.ds[0]);
    mbar_wait(&sm.s_full, 0, 74);
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = ch * 2 + cc;
      uint32_t s[32], d[32];
      tmem_ld_x32(tmem + lane_addr + c * 32, s);
      tmem_ld_x32(tmem + lane_addr + C_DP + c * 32, d);
      tc_wait_ld();
      const int key0 = c * 32;
      uint32_t pp[16], dd[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float pv[2], dv[2];
        uint32_t bits = 0xFFFFFFFFu;
        if (dp.thr16) bits = drop_bits(drop_seed, dp.site, static_cast<uint32_t>(hrow + r), (key0 >> 1) + i);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kk = 2 * i + e;
          float p = ex2_approx(fmaf(__uint_as_float(s[kk]), scale_log2, -l2));
          p = (key0 + kk < Q) ? p : 0.f;
          const float keep = (((bits >> (16 * e)) & 0xFFFFu) >= dp.thr16) ? drop_s : 0.f;
          pv[e] = p * p_scale * keep;
          dv[e] = p * (__uint_as_float(d[kk]) * keep * dp_scale - dl) * ds_scale;
        }
        pp[i] = pack_bf16x2(pv[0], pv[1]);
        dd[i] = pack_bf16x2(dv[0], dv[1]);
      }
      // row r of the tile, columns [32c, 32c+32): four 16-byte units of the 64-column chunk c/2, SW128-swizzled
      const uint32_t chunk = (c >> 1) * CHUNK_BYTES;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t off = chunk + swz_offset<128>(r, (c & 1) * 4 + u);
        sts_u4(a_p + off, pp[4 * u], pp[4 * u + 1], pp[4 * u + 2], pp[4 * u + 3]);
        sts_u4(a_ds + off, dd[4 * u], dd[4 * u + 1], dd[4 * u + 2], dd[4 * u + 3]);
      }
    }
    fence_proxy_async_smem();  // the tiles were written with ordinary stores; the tensor core reads them
    tc_fence_before();
    mbar_arrive(&sm.p_full);
    mbar_wait(&sm.g_full, 0, 75);
    tc_fence_after();
    // epilogue: thread = row of dQ (query) / dK, dV (key); head-major outputs, D contiguous values per row
    const size_t orow = (static_cast<size_t>(hrow) + r) * D;
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const uint32_t col = (w == 0 ? C_DQ : (w == 1 ? C_DK : C_DV));
      __nv_bfloat16* dst = (w == 0 ? dq_out : (w == 1 ? dk_out : dv_out)) + orow;
#pragma unroll
      for (int cc = 0; cc < D / 64; ++cc) {
        const int c = ch * (D / 64) + cc;  // this thread's half of the row
        uint32_t v[32];
        tmem_ld_x32(tmem + lane_addr + col + c * 32, v);
        tc_wait_ld();
        if (valid) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[8 * u]), __uint_as_float(v[8 * u + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[8 * u + 2]), __uint_as_float(v[8 * u + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[8 * u + 4]), __uint_as_float(v[8 * u + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[8 * u + 6]), __uint_as_float(v[8 * u + 7]));
            reinterpret_cast<uint4*>(dst + c * 32)[u] = o;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(NTHREADS_F, 1)
dec_attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tq1, const __grid_constant__ CUtensorMap tk1,
                          const __grid_constant__ CUtensorMap tv1, const __grid_constant__ CUtensorMap td1,
                          const __grid_constant__ CUtensorMap tq2, const __grid_constant__ CUtensorMap tk2,
                          const __grid_constant__ CUtensorMap tv2, const __grid_constant__ CUtensorMap td2, ParamsF p) {
  extern __shared__ uint8_t smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8 && lane == 0) {
    mbar_init(&sm.qk_full, 1);
    mbar_init(&sm.dov_full, 1);
    mbar_init(&sm.s_full, 1);
    mbar_init(&sm.p_full, 256);
    mbar_init(&sm.g_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const int hy = blockIdx.x, b = blockIdx.y;
  const float log2e = 1.4426950408889634f, r = 0.08838834764831845f;
  if (hy < 8)
    body_fused<64>(sm, &tq1, &tk1, &tv1, &td1, p.lse1, p.delta1, p.dqkv, p.dqkv + p.rows * 512, p.dqkv + 2 * p.rows * 512,
                   p.Q, hy, b, log2e * 0.125f, 1.f, 1.f, 0.125f, tmem, p.dp);
  else
    body_fused<128>(sm, &tq2, &tk2, &tv2, &td2, p.lse2, p.delta2, p.dcat, p.dcat + p.rows * 1024,
                    p.dcat + 2 * p.rows * 1024, p.Q, hy - 8, b, log2e, r, r, 1.f, tmem, Drop{nullptr, 0u, 0u});
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// Backward of dec_qkv_prep: gather formulation of the scatter-add (no atomics, deterministic).
//   d_x_w[b,i] = d_self_w[b,i] + sum_{i': L_i' = i} d_cat_w[b,i'][left] + sum_{i': R_i' = i} d_cat_w[b,i'][right]
//   d_qkv_obj = [d_q | d_k | d_v] token-major;  d_qk_pos = [d_q lo+hi halves | d_k lo+hi halves]
// One block per (b, i), 64 threads x 8 channels; the image's pairs sit in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void dec_qkv_prep_bwd_kernel(const __nv_bfloat16* __restrict__ d_qkv,  // [3][B,8,Q,64]
                                        const __nv_bfloat16* __restrict__ d_cat,  // [3][B,8,Q,128]
                                        const int32_t* __restrict__ pairs, __nv_bfloat16* __restrict__ d_obj,
                                        __nv_bfloat16* __restrict__ d_pos, int ld_pos, int Q, int rows) {
  extern __shared__ int32_t sp[];  // [Q][2]
  __shared__ float fold[2][512];
  const int row = blockIdx.x;
  const int b = row / Q, i = row - b * Q;
  const int t = threadIdx.x;
  for (int k = t; k < 2 * Q; k += blockDim.x) sp[k] = pairs[static_cast<size_t>(b) * Q * 2 + k];
  __syncthreads();
  const int hh = t >> 3, within = (t & 7) * 8;
  const size_t hm_base = static_cast<size_t>(b * 8 + hh) * Q;
  const size_t self_stride = static_cast<size_t>(rows) * 512, cat_stride = static_cast<size_t>(rows) * 1024;
  float acc[3][8];
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    const uint4 u = *reinterpret_cast<const uint4*>(d_qkv + w * self_stride + (hm_base + i) * 64 + within);
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[w][2 * e] = __uint_as_float(uw[e] << 16);
      acc[w][2 * e + 1] = __uint_as_float(uw[e] & 0xffff0000u);
    }
  }
  // Which (i', side) gather from this row: key k = 2 i' + side matches iff sp[k] == i.  Warp 0 compacts the matching
  // keys in ASCENDING order with ballots (7 rounds at Q = 100) instead of every thread scanning all Q pairs -- the
  // scan was most of this kernel's time -- and everybody then adds the few (two on average) matches in that order,
  // i.e. in the order of the sequential scan: same rounding, still no atomics.
  __shared__ int32_t hits[256];  // (more than 256 gathers into one row: the tail is handled by the scan below)
  __shared__ int32_t n_hits;
  if (t < 32) {
    int n = 0;
    for (int base = 0; base < 2 * Q; base += 32) {
      const int k = base + t;
      const unsigned m = __ballot_sync(0xffffffffu, k < 2 * Q && sp[k] == i);
      if (m & (1u << t)) {
        const int pos = n + __popc(m & ((1u << t) - 1u));
        if (pos < 256) hits[pos] = k;
      }
      n += __popc(m);
    }
    if (t == 0) n_hits = n;
  }
  __syncthreads();
  const int nh = n_hits;
  auto gather = [&](int k) {
    const int ip = k >> 1, side = k & 1;
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const uint4 u = *reinterpret_cast<const uint4*>(d_cat + w * cat_stride + (hm_base + ip) * 128 + side * 64 + within);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[w][2 * e] += __uint_as_float(uw[e] << 16);
        acc[w][2 * e + 1] += __uint_as_float(uw[e] & 0xffff0000u);
      }
    }
  };
  for (int h = 0; h < (nh < 256 ? nh : 256); ++h) gather(hits[h]);
  if (nh > 256)  // (degenerate pairing: keep the order, scan for the rest)
    for (int k = hits[255] + 1; k < 2 * Q; ++k)
      if (sp[k] == i) gather(k);
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(acc[w][2 * e], acc[w][2 * e + 1]);
    *reinterpret_cast<uint4*>(d_obj + static_cast<size_t>(row) * 1536 + w * 512 + t * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    if (w < 2) {
#pragma unroll
      for (int e = 0; e < 8; ++e) fold[w][t * 8 + e] = acc[w][e];
    }
  }
  __syncthreads();
  // q_pos / k_pos were added to BOTH 256-wide halves (decoder_block.py:169,173): fold the halves
  for (int k = t; k < 512; k += blockDim.x) {
    const int w = k >> 8, c = k & 255;
    d_pos[static_cast<size_t>(row) * ld_pos + w * 256 + c] = __float2bfloat16(fold[w][c] + fold[w][256 + c]);
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_dec_qkv_prep_bwd(const void* d_qkv, const void* d_cat, const int32_t* pairs, void* d_qkv_obj,
                                      void* d_qk_pos, int ld_pos, int B, int Q, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(d_qkv && d_cat && pairs && d_qkv_obj && d_qk_pos && B > 0 && Q > 0, "null pointer / shape");
  DESTR_CHECK_ARG(ld_pos >= 512, "ld_pos");
  dec_qkv_prep_bwd_kernel<<<B * Q, 64, 2 * Q * sizeof(int32_t), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(d_qkv), static_cast<const __nv_bfloat16*>(d_cat), pairs,
      static_cast<__nv_bfloat16*>(d_qkv_obj), static_cast<__nv_bfloat16*>(d_qk_pos), ld_pos, Q, B * Q);
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_dec_self_pair_attn_bwd_ds(const void* qkv, const void* cat, const void* do1, const void* do2,
                                               const float* lse1, const float* lse2, const float* delta1,
                                               const float* delta2, void* P1, void* dS1, void* P2, void* dS2, int B,
                                               int Q, const uint32_t* drop_seed, uint32_t drop_thr16,
                                               uint32_t drop_site, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(qkv && cat && do1 && do2 && lse1 && lse2 && delta1 && delta2 && P1 && dS1 && P2 && dS2,
                  "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && Q <= BT * MAX_TILES, "Q must be <= 384");
  const uint64_t rows = static_cast<uint64_t>(B) * Q;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* c = static_cast<const __nv_bfloat16*>(cat);
  CUtensorMap t[8];
  int rc;
  for (int w = 0; w < 3; ++w) {
    if ((rc = make_tmap_bf16_2d(&t[w], x + w * rows * 512, rows * 8, 64, 64, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
    if ((rc = make_tmap_bf16_2d(&t[4 + w], c + w * rows * 1024, rows * 8, 128, 128, BT, 64,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  if ((rc = make_tmap_bf16_2d(&t[3], do1, rows * 8, 64, 64, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&t[7], do2, rows * 8, 128, 128, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const size_t smem = sizeof(Smem) + 1024;
  DESTR_SMEM_OPTIN(dec_attn_bwd_ds_kernel, smem);
  const int Qp = ceil_div(Q, BT) * BT;
  Params p{lse1, lse2, delta1, delta2, static_cast<__nv_bfloat16*>(P1), static_cast<__nv_bfloat16*>(dS1),
           static_cast<__nv_bfloat16*>(P2), static_cast<__nv_bfloat16*>(dS2), Q, Qp,
           Drop{drop_seed, drop_thr16, drop_site}};
  dim3 grid(ceil_div(Q, BT), 16, B);
  dec_attn_bwd_ds_kernel<<<grid, NTHREADS, smem, static_cast<cudaStream_t>(stream)>>>(t[0], t[1], t[2], t[3], t[4],
                                                                                      t[5], t[6], t[7], p);
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_dec_self_pair_attn_bwd(const void* qkv, const void* cat, const void* do1, const void* do2,
                                            const float* lse1, const float* lse2, const float* delta1,
                                            const float* delta2, void* d_qkv, void* d_cat, int B, int Q,
                                            const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site,
                                            void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(qkv && cat && do1 && do2 && lse1 && lse2 && delta1 && delta2 && d_qkv && d_cat, "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && Q <= BT, "fused backward: Q must be <= 128 (use the _ds entry point above)");
  const uint64_t rows = static_cast<uint64_t>(B) * Q;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* c = static_cast<const __nv_bfloat16*>(cat);
  CUtensorMap t[8];
  int rc;
  for (int w = 0; w < 3; ++w) {
    if ((rc = make_tmap_bf16_2d(&t[w], x + w * rows * 512, rows * 8, 64, 64, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
    if ((rc = make_tmap_bf16_2d(&t[4 + w], c + w * rows * 1024, rows * 8, 128, 128, BT, 64,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  if ((rc = make_tmap_bf16_2d(&t[3], do1, rows * 8, 64, 64, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&t[7], do2, rows * 8, 128, 128, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const size_t smem = sizeof(SmemF) + 1024;
  DESTR_SMEM_OPTIN(dec_attn_bwd_fused_kernel, smem);
  ParamsF p{lse1, lse2, delta1, delta2, static_cast<__nv_bfloat16*>(d_qkv), static_cast<__nv_bfloat16*>(d_cat), Q,
            static_cast<size_t>(rows), Drop{drop_seed, drop_thr16, drop_site}};
  dec_attn_bwd_fused_kernel<<<dim3(16, B), NTHREADS_F, smem, static_cast<cudaStream_t>(stream)>>>(
      t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], p);
  DESTR_LAUNCH_CHECK();
  return 0;
}
