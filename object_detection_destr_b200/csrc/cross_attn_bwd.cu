// Split cross-attention backward, stage 1 (tcgen05 / TMEM / TMA): recompute the scores and dP on the
// tensor cores and apply the softmax backward, emitting P and dS in bf16.
//
// Forward (cross_attn_fwd.cu; reference decoder_block.py:189-217, 246-251 -> self_attention.py:26-45):
//   S_br = ([q_obj_br | q_pos] . [k_enc | k_pos]^T) * scale + mask,  P_br = softmax(S_br),  O_br = P_br . V
// Backward per (b, branch br, query q, key k), with lse and delta = rowsum(dO o O) from the forward:
//   P  = 2^(c*S - lse)        dP = dO_br . V^T        dS = P o (dP - delta) * scale
// The decoder's score matrices are small (2Q x N per image), so unlike the encoder (N x N, fully fused
// in enc_attn_bwd.cu) P and dS ARE written out -- 2 x 3.4 MB at config 2 -- and the five contractions
//   dV = P^T dO,  dK_enc = dS^T q_obj,  dK_pos = dSsum^T q_pos,  dq_obj = dS k_enc,  dq_pos = dSsum k_pos
// run as plain batched GEMMs.  Rows are ordered (2q + br) so that q_obj [B*Q,512] viewed as [B,2Q,256],
// dO likewise, and the dq_obj result are all plain views; dSsum = dS_cls + dS_reg is emitted here too.
//
// grid = (key tiles, query tiles, B); one CTA handles BOTH branches of its (query tile, key tile):
//   24 chunk pairs (64-wide, SW128, K-major) stream through a 4-stage TMA ring:
//     per branch: 4 x (q_obj_br, k_enc), 4 x (q_pos, k_pos) -> S_br ;  4 x (dO_br, V) -> dP_br
//   TMEM: S_0 [0,128) dP_0 [128,256) S_1 [256,384) dP_1 [384,512)  (branch 1's MMAs overlap branch 0's math)
// warps 0-3 math (thread <-> query row <-> TMEM lane), warp 4 TMA, warp 5 MMA.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {
namespace {

constexpr int BT = 128;
constexpr int NSTAGE = 4;
constexpr int NTHREADS = 192;
constexpr uint32_t CHUNK_BYTES = BT * 128;

struct __align__(1024) Smem {
  uint8_t a[NSTAGE][CHUNK_BYTES];
  uint8_t b[NSTAGE][CHUNK_BYTES];
  uint64_t full[NSTAGE];
  uint64_t empty[NSTAGE];
  uint64_t sdp_full[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(NTHREADS, 1)
cross_attn_bwd_ds_kernel(const __grid_constant__ CUtensorMap tm_qobj, const __grid_constant__ CUtensorMap tm_qpos,
                         const __grid_constant__ CUtensorMap tm_kenc, const __grid_constant__ CUtensorMap tm_kpos,
                         const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                         const uint32_t* __restrict__ mask_bits, int words_per_row, const float* __restrict__ lse,
                         const float* __restrict__ delta, __nv_bfloat16* __restrict__ P_all,
                         __nv_bfloat16* __restrict__ dS_all, __nv_bfloat16* __restrict__ dS_sum, int Q, int N, int Np,
                         float scale, float scale_log2, Drop dp) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;
  const int qrow0 = b * Q + qt * BT;
  const int krow0 = b * N + j * BT;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.sdp_full[0], 1);
    mbar_init(&sm.sdp_full[1], 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 4) {
    if (elect_one()) {
      for (int t = 0; t < 24; ++t) {
        const int br = t / 12, c = t % 12;
        const int s = t % NSTAGE;
        mbar_wait(&sm.empty[s], ((t / NSTAGE) & 1) ^ 1, 41);
        mbar_arrive_expect_tx(&sm.full[s], 2 * CHUNK_BYTES);
        if (c < 4) {
          tma_load_2d(sm.a[s], &tm_qobj, &sm.full[s], br * 256 + c * 64, qrow0);
          tma_load_2d(sm.b[s], &tm_kenc, &sm.full[s], c * 64, krow0);
        } else if (c < 8) {
          tma_load_2d(sm.a[s], &tm_qpos, &sm.full[s], (c - 4) * 64, qrow0);
          tma_load_2d(sm.b[s], &tm_kpos, &sm.full[s], (c - 4) * 64, krow0);
        } else {
          tma_load_2d(sm.a[s], &tm_do, &sm.full[s], br * 256 + (c - 8) * 64, qrow0);
          tma_load_2d(sm.b[s], &tm_v, &sm.full[s], (c - 8) * 64, krow0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BT, BT, false, false);
      for (int t = 0; t < 24; ++t) {
        const int br = t / 12, c = t % 12;
        const int s = t % NSTAGE;
        mbar_wait(&sm.full[s], (t / NSTAGE) & 1, 42);
        tc_fence_after();
        const uint32_t dst = tmem + br * 256 + (c < 8 ? 0 : 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_ss(dst, umma_smem_desc(smem_u32(sm.a[s]) + ks * 32, 16, 1024, SWZ_128B),
                  umma_smem_desc(smem_u32(sm.b[s]) + ks * 32, 16, 1024, SWZ_128B), idesc,
                  ((c != 0 && c != 8) || ks > 0) ? 1u : 0u);
        }
        tc_commit(&sm.empty[s]);
        if (c == 11) tc_commit(&sm.sdp_full[br]);
      }
    }
    __syncwarp();
  } else {
    const int wq = warp;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const int q = qt * BT + wq * 32 + lane;
    const bool qvalid = q < Q;
    const uint32_t drop_seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
    const float drop_s = drop_scale(dp.thr16);
    const uint4 mw = *reinterpret_cast<const uint4*>(mask_bits + static_cast<size_t>(b) * words_per_row + j * 4);
    const uint32_t mwa[4] = {mw.x, mw.y, mw.z, mw.w};
    uint32_t stash[4][16];  // packed dS of branch 0
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      const float l2 = qvalid ? lse[(static_cast<size_t>(b) * 2 + br) * Q + q] : INFINITY;
      const float dl = qvalid ? delta[(static_cast<size_t>(b) * 2 + br) * Q + q] : 0.f;
      mbar_wait(&sm.sdp_full[br], 0, 43);
      tc_fence_after();
      const size_t row = (static_cast<size_t>(b) * Q + (qvalid ? q : 0)) * 2 + br;
      __nv_bfloat16* prow = P_all + row * Np + j * BT;
      __nv_bfloat16* drow = dS_all + row * Np + j * BT;
      __nv_bfloat16* srow = dS_sum + (static_cast<size_t>(b) * Q + (qvalid ? q : 0)) * Np + j * BT;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t s[32], d[32];
        tmem_ld_x32(tmem + lane_addr + br * 256 + c * 32, s);
        tmem_ld_x32(tmem + lane_addr + br * 256 + 128 + c * 32, d);
        tc_wait_ld();
        uint32_t pp[16], dd[16], ss[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pv[2], dv[2];
          uint32_t bits = 0xFFFFFFFFu;
          if (dp.thr16)
            bits = drop_bits(drop_seed, dp.site, static_cast<uint32_t>((b * 2 + br) * Q + q), j * (BT / 2) + c * 16 + i);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = 2 * i + e;
            float p = ex2_approx(fmaf(__uint_as_float(s[k]), scale_log2, -l2));
            p = ((mwa[c] >> k) & 1u) ? 0.f : p;
            // forward: O = dropout(P) V  ->  dV takes the dropped P, dP passes through the same mask
            const float keep = (((bits >> (16 * e)) & 0xFFFFu) >= dp.thr16) ? drop_s : 0.f;
            pv[e] = p * keep;
            dv[e] = p * (__uint_as_float(d[k]) * keep - dl) * scale;
          }
          pp[i] = pack_bf16x2(pv[0], pv[1]);
          dd[i] = pack_bf16x2(dv[0], dv[1]);
          if (br == 0) {
            stash[c][i] = dd[i];
          } else {
            const uint32_t o = stash[c][i];
            ss[i] = pack_bf16x2(dv[0] + __uint_as_float(o << 16), dv[1] + __uint_as_float(o & 0xffff0000u));
          }
        }
        if (qvalid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            reinterpret_cast<uint4*>(prow + c * 32)[i] = make_uint4(pp[4 * i], pp[4 * i + 1], pp[4 * i + 2], pp[4 * i + 3]);
            reinterpret_cast<uint4*>(drow + c * 32)[i] = make_uint4(dd[4 * i], dd[4 * i + 1], dd[4 * i + 2], dd[4 * i + 3]);
            if (br == 1)
              reinterpret_cast<uint4*>(srow + c * 32)[i] =
                  make_uint4(ss[4 * i], ss[4 * i + 1], ss[4 * i + 2], ss[4 * i + 3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// FUSED backward for Q <= 128 (one query tile, the training shape): CTA = (key tile j, image b), both branches.
//   phase 1 (as above)   S_br, dP_br on the tensor cores -> softmax backward -> P_br, dS_br as bf16 TILES IN SHARED
//                        MEMORY ([128 queries][128 keys], two SW128 chunks of 64 keys each); P never goes to HBM
//   phase 2              dV_j  = sum_br P_br^T  dO_br        TMEM [0,256)      A = tile read MN-major (contraction
//                        dKe_j = sum_br dS_br^T q_obj_br     TMEM [256,512)        over the 128 query rows),
//   phase 3              dKp_j = sum_br dS_br^T q_pos        TMEM [0,256)      B = [128 queries][64] chunks, MN-major
//   The key-side gradients of a key tile are complete inside its CTA: they are stored (bf16) straight into the packed
//   d_kv_all / d_kpos_all slices -- no atomics, no partials.  The query side (dq_obj = dS k_enc, dq_pos =
//   (dS_cls + dS_reg) k_pos) contracts over ALL keys of the image: the kernel also writes dS and dS_cls + dS_reg
//   (bf16, 2 x 3.4 MB + 1.7 MB at config 2) and two batched tcgen05 GEMMs (destr_gemm_bf16_batched) finish.
// TMA ring: 3 stages of (A 16 KB, B 16 KB); phases 2-3 stream single B chunks through the same ring.
// ------------------------------------------------------------------------------------------------
constexpr int NSTAGE_F = 3;
constexpr int NTHREADS_F = 320;  // 8 math warps + TMA warp + MMA warp
struct __align__(1024) SmemF {
  uint8_t a[NSTAGE_F][CHUNK_BYTES];
  uint8_t b[NSTAGE_F][CHUNK_BYTES];
  uint8_t p[2][2][CHUNK_BYTES];   // [branch][64-key chunk]
  uint8_t ds[2][2][CHUNK_BYTES];
  uint64_t full[NSTAGE_F];
  uint64_t empty[NSTAGE_F];
  uint64_t sdp_full[2];
  uint64_t tiles_full;   // P / dS tiles written, S / dP read out of TMEM (128 arrivals)
  uint64_t g_full[2];    // phase 2 / phase 3 accumulators complete
  uint64_t g_drained;    // phase 2 accumulators read out (128 arrivals)
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(NTHREADS_F, 1)
cross_attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tm_qobj, const __grid_constant__ CUtensorMap tm_qpos,
                            const __grid_constant__ CUtensorMap tm_kenc, const __grid_constant__ CUtensorMap tm_kpos,
                            const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                            const uint32_t* __restrict__ mask_bits, int words_per_row, const float* __restrict__ lse,
                            const float* __restrict__ delta, __nv_bfloat16* __restrict__ dS_all,
                            __nv_bfloat16* __restrict__ dS_sum, __nv_bfloat16* __restrict__ dke, int ld_dke,
                            __nv_bfloat16* __restrict__ dkp, int ld_dkp, __nv_bfloat16* __restrict__ dv, int ld_dv,
                            int Q, int N, int Np, float scale, float scale_log2, Drop dp) {
  extern __shared__ uint8_t smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, b = blockIdx.y;
  const int qrow0 = b * Q;
  const int krow0 = b * N + j * BT;
  constexpr int NS = NSTAGE_F;
  constexpr int T1 = 24, T2 = 16, T3 = 4;  // ring entries of the three phases

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.sdp_full[0], 1);
    mbar_init(&sm.sdp_full[1], 1);
    mbar_init(&sm.tiles_full, 256);
    mbar_init(&sm.g_full[0], 1);
    mbar_init(&sm.g_full[1], 1);
    mbar_init(&sm.g_drained, 256);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 8) {
    if (elect_one()) {
      for (int t = 0; t < T1 + T2 + T3; ++t) {
        const int s = t % NS;
        mbar_wait(&sm.empty[s], ((t / NS) & 1) ^ 1, 41);
        if (t < T1) {
          const int br = t / 12, c = t % 12;
          mbar_arrive_expect_tx(&sm.full[s], 2 * CHUNK_BYTES);
          if (c < 4) {
            tma_load_2d(sm.a[s], &tm_qobj, &sm.full[s], br * 256 + c * 64, qrow0);
            tma_load_2d(sm.b[s], &tm_kenc, &sm.full[s], c * 64, krow0);
          } else if (c < 8) {
            tma_load_2d(sm.a[s], &tm_qpos, &sm.full[s], (c - 4) * 64, qrow0);
            tma_load_2d(sm.b[s], &tm_kpos, &sm.full[s], (c - 4) * 64, krow0);
          } else {
            tma_load_2d(sm.a[s], &tm_do, &sm.full[s], br * 256 + (c - 8) * 64, qrow0);
            tma_load_2d(sm.b[s], &tm_v, &sm.full[s], (c - 8) * 64, krow0);
          }
        } else {
          // phases 2 / 3: one [128 queries][64 columns] chunk, the MN-major B operand of a key-side product
          const int u = t - T1;
          mbar_arrive_expect_tx(&sm.full[s], CHUNK_BYTES);
          if (u < 8) {            // dO_br chunk c        (dV)
            tma_load_2d(sm.a[s], &tm_do, &sm.full[s], (u >> 2) * 256 + (u & 3) * 64, qrow0);
          } else if (u < 16) {    // q_obj_br chunk c     (dK_enc)
            tma_load_2d(sm.a[s], &tm_qobj, &sm.full[s], ((u - 8) >> 2) * 256 + (u & 3) * 64, qrow0);
          } else {                // q_pos chunk c        (dK_pos)
            tma_load_2d(sm.a[s], &tm_qpos, &sm.full[s], (u - 16) * 64, qrow0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BT, BT, false, false);
      constexpr uint32_t id_nn = umma_idesc_bf16(BT, 64, true, true);     // A MN-major (tile), B MN-major, N = 64
      constexpr uint64_t D_BMN = umma_desc_const(16, 1024, SWZ_128B);       // one 64-wide atom, k-step 2048 B
      constexpr uint64_t D_AMN = umma_desc_const(16384, 1024, SWZ_128B);    // two 64-wide atoms 16 KB apart
      int t = 0;
      for (; t < T1; ++t) {
        const int br = t / 12, c = t % 12;
        const int s = t % NS;
        mbar_wait(&sm.full[s], (t / NS) & 1, 42);
        tc_fence_after();
        const uint32_t dst = tmem + br * 256 + (c < 8 ? 0 : 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_ss(dst, umma_smem_desc(smem_u32(sm.a[s]) + ks * 32, 16, 1024, SWZ_128B),
                  umma_smem_desc(smem_u32(sm.b[s]) + ks * 32, 16, 1024, SWZ_128B), idesc,
                  ((c != 0 && c != 8) || ks > 0) ? 1u : 0u);
        }
        tc_commit(&sm.empty[s]);
        if (c == 11) tc_commit(&sm.sdp_full[br]);
      }
      mbar_wait(&sm.tiles_full, 0, 44);  // tiles in shared memory, S / dP consumed: TMEM is free for the accumulators
      tc_fence_after();
      for (; t < T1 + T2 + T3; ++t) {
        const int u = t - T1;
        const int s = t % NS;
        mbar_wait(&sm.full[s], (t / NS) & 1, 45);
        if (u == T2) {  // phase 3 reuses dV's columns
          mbar_wait(&sm.g_drained, 0, 46);
        }
        tc_fence_after();
        const uint64_t db = D_BMN + (smem_u32(sm.a[s]) >> 4);
        if (u < T2) {
          const int br = (u >> 2) & 1, c = u & 3;
          const uint32_t a_tile = (u < 8) ? smem_u32(sm.p[br][0]) : smem_u32(sm.ds[br][0]);
          const uint32_t dst = tmem + (u < 8 ? 0 : 256) + c * 64;
#pragma unroll
          for (int ks = 0; ks < BT / 16; ++ks)  // contraction over the 128 query rows
            umma_ss(dst, D_AMN + ((a_tile + ks * 2048) >> 4), db + ks * 128, id_nn, (br > 0 || ks > 0) ? 1u : 0u);
        } else {
          const int c = u - T2;
#pragma unroll
          for (int br = 0; br < 2; ++br)
#pragma unroll
            for (int ks = 0; ks < BT / 16; ++ks)
              umma_ss(tmem + c * 64, D_AMN + ((smem_u32(sm.ds[br][0]) + ks * 2048) >> 4), db + ks * 128, id_nn,
                      (br > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&sm.empty[s]);
        if (u == T2 - 1) tc_commit(&sm.g_full[0]);
      }
      tc_commit(&sm.g_full[1]);
    }
    __syncwarp();
  } else {
    // 8 math warps: warps w and w + 4 own the same 32 TMEM lanes (rows) and split the columns, so the per-thread
    // instruction stream of the softmax backward and of the drains is half as long (the kernel is a latency chain)
    const int wq = warp & 3, ch = warp >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const int r = wq * 32 + lane;  // query row (phase 1) / key row of the tile (drains)
    const int q = r;
    const bool qvalid = q < Q;
    const uint32_t drop_seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
    const float drop_s = drop_scale(dp.thr16);
    const uint4 mw = *reinterpret_cast<const uint4*>(mask_bits + static_cast<size_t>(b) * words_per_row + j * 4);
    uint32_t stash[2][16];  // packed dS of branch 0 (this thread's two 32-key chunks)
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      const float l2 = qvalid ? lse[(static_cast<size_t>(b) * 2 + br) * Q + q] : INFINITY;
      const float dl = qvalid ? delta[(static_cast<size_t>(b) * 2 + br) * Q + q] : 0.f;
      mbar_wait(&sm.sdp_full[br], 0, 43);
      tc_fence_after();
      const size_t row = (static_cast<size_t>(b) * Q + (qvalid ? q : 0)) * 2 + br;
      __nv_bfloat16* drow = dS_all + row * Np + j * BT;
      __nv_bfloat16* srow = dS_sum + (static_cast<size_t>(b) * Q + (qvalid ? q : 0)) * Np + j * BT;
      const uint32_t a_p = smem_u32(sm.p[br][0]), a_ds = smem_u32(sm.ds[br][0]);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = ch * 2 + cc;  // 32-key chunk of the tile
        const uint32_t mword = ch == 0 ? (cc == 0 ? mw.x : mw.y) : (cc == 0 ? mw.z : mw.w);
        uint32_t s[32], d[32];
        tmem_ld_x32(tmem + lane_addr + br * 256 + c * 32, s);
        tmem_ld_x32(tmem + lane_addr + br * 256 + 128 + c * 32, d);
        tc_wait_ld();
        uint32_t pp[16], dd[16], ss[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pv[2], dv_[2];
          uint32_t bits = 0xFFFFFFFFu;
          if (dp.thr16)
            bits = drop_bits(drop_seed, dp.site, static_cast<uint32_t>((b * 2 + br) * Q + q), j * (BT / 2) + c * 16 + i);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = 2 * i + e;
            float pr = ex2_approx(fmaf(__uint_as_float(s[k]), scale_log2, -l2));
            pr = ((mword >> k) & 1u) ? 0.f : pr;
            const float keep = (((bits >> (16 * e)) & 0xFFFFu) >= dp.thr16) ? drop_s : 0.f;
            pv[e] = pr * keep;
            dv_[e] = pr * (__uint_as_float(d[k]) * keep - dl) * scale;
          }
          pp[i] = pack_bf16x2(pv[0], pv[1]);
          dd[i] = pack_bf16x2(dv_[0], dv_[1]);
          if (br == 0) {
            stash[cc][i] = dd[i];
          } else {
            const uint32_t o = stash[cc][i];
            ss[i] = pack_bf16x2(dv_[0] + __uint_as_float(o << 16), dv_[1] + __uint_as_float(o & 0xffff0000u));
          }
        }
        // tiles for the key-side products: row r, columns [32c, 32c+32) of chunk c/2 (rows >= Q are zero: lse = inf)
        const uint32_t chunk = (c >> 1) * CHUNK_BYTES;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t off = chunk + swz_offset<128>(r, (c & 1) * 4 + u);
          sts_u4(a_p + off, pp[4 * u], pp[4 * u + 1], pp[4 * u + 2], pp[4 * u + 3]);
          sts_u4(a_ds + off, dd[4 * u], dd[4 * u + 1], dd[4 * u + 2], dd[4 * u + 3]);
        }
        if (qvalid) {  // dS for the query-side GEMMs
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            reinterpret_cast<uint4*>(drow + c * 32)[i] = make_uint4(dd[4 * i], dd[4 * i + 1], dd[4 * i + 2], dd[4 * i + 3]);
            if (br == 1)
              reinterpret_cast<uint4*>(srow + c * 32)[i] =
                  make_uint4(ss[4 * i], ss[4 * i + 1], ss[4 * i + 2], ss[4 * i + 3]);
          }
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(&sm.tiles_full);
    // ---- drains: thread = key row of the tile ----
    const int key = j * BT + r;
    const bool kvalid = key < N;
    const size_t grow = static_cast<size_t>(b) * N + (kvalid ? key : 0);
    auto drain = [&](uint32_t col, __nv_bfloat16* dst) {  // this thread: 128 of the row's 256 columns
#pragma unroll
      for (int c = ch * 4; c < ch * 4 + 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem + lane_addr + col + c * 32, v);
        tc_wait_ld();
        if (kvalid) {
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
    };
    mbar_wait(&sm.g_full[0], 0, 47);
    tc_fence_after();
    drain(0, dv + grow * ld_dv);
    tc_fence_before();
    mbar_arrive(&sm.g_drained);  // dV's columns may be overwritten by phase 3
    drain(256, dke + grow * ld_dke);
    mbar_wait(&sm.g_full[1], 0, 48);
    tc_fence_after();
    drain(0, dkp + grow * ld_dkp);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// delta[b, br, q] = sum_c dO[b*Q+q, br*256+c] * O[b*Q+q, br*256+c]   (one warp per query row)
__global__ void cross_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                   float* __restrict__ delta, int B, int Q) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B * Q) return;
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {  // lane owns 16 channels: [lane*16, lane*16+16)
    const uint4 a = *reinterpret_cast<const uint4*>(o + static_cast<size_t>(row) * 512 + lane * 16 + u * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(d_o + static_cast<size_t>(row) * 512 + lane * 16 + u * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(gw[i] << 16), s);
      s = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(gw[i] & 0xffff0000u), s);
    }
  }
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);  // within 16-lane halves
  if ((lane & 15) == 0) {
    const int b = row / Q, q = row - b * Q, br = lane >> 4;
    delta[(static_cast<size_t>(b) * 2 + br) * Q + q] = s;
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_split_cross_attn_bwd_ds(const void* q_obj, const void* q_pos, const void* k_enc,
                                             const void* k_pos, const void* v, int ld_kenc, int ld_kpos, int ld_v,
                                             const uint32_t* mask_bits, int words_per_row, const void* out,
                                             const void* dout, const float* lse, float* delta, void* P_all,
                                             void* dS_all, void* dS_sum, int B, int Q, int N, float scale,
                                             const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site,
                                             void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(q_obj && q_pos && k_enc && k_pos && v && mask_bits && out && dout && lse && delta && P_all &&
                      dS_all && dS_sum,
                  "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && N > 0, "shape");
  const int nqt = ceil_div(Q, BT), nkv = ceil_div(N, BT);
  DESTR_CHECK_ARG(words_per_row >= nkv * 4, "words_per_row");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t qrows = static_cast<uint64_t>(B) * Q, krows = static_cast<uint64_t>(B) * N;
  CUtensorMap tqo, tqp, tke, tkp, tv, tdo;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tqo, q_obj, qrows, 512, 512, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tqp, q_pos, qrows, 256, 256, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tke, k_enc, krows, 256, ld_kenc, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tkp, k_pos, krows, 256, ld_kpos, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, v, krows, 256, ld_v, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, dout, qrows, 512, 512, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const size_t smem = sizeof(Smem) + 1024;
  DESTR_SMEM_OPTIN(cross_attn_bwd_ds_kernel, smem);
  cross_delta_kernel<<<ceil_div((int)qrows, 8), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out),
                                                              static_cast<const __nv_bfloat16*>(dout), delta, B, Q);
  DESTR_LAUNCH_CHECK();
  dim3 grid(nkv, nqt, B);
  cross_attn_bwd_ds_kernel<<<grid, NTHREADS, smem, st>>>(
      tqo, tqp, tke, tkp, tv, tdo, mask_bits, words_per_row, lse, delta, static_cast<__nv_bfloat16*>(P_all),
      static_cast<__nv_bfloat16*>(dS_all), static_cast<__nv_bfloat16*>(dS_sum), Q, N, nkv * BT, scale,
      scale * 1.4426950408889634f, Drop{drop_seed, drop_thr16, drop_site});
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_split_cross_attn_bwd_fused(const void* q_obj, const void* q_pos, const void* k_enc,
                                                const void* k_pos, const void* v, int ld_kenc, int ld_kpos, int ld_v,
                                                const uint32_t* mask_bits, int words_per_row, const void* out,
                                                const void* dout, const float* lse, float* delta, void* dS_all,
                                                void* dS_sum, void* dk_enc, int ld_dke, void* dk_pos, int ld_dkp,
                                                void* dv, int ld_dv, int B, int Q, int N, float scale,
                                                const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site,
                                                void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(q_obj && q_pos && k_enc && k_pos && v && mask_bits && out && dout && lse && delta && dS_all &&
                      dS_sum && dk_enc && dk_pos && dv,
                  "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && Q <= BT && N > 0, "fused backward: Q must be <= 128 (use the _ds entry point)");
  DESTR_CHECK_ARG(ld_dke % 8 == 0 && ld_dkp % 8 == 0 && ld_dv % 8 == 0 && ld_dke >= 256 && ld_dkp >= 256 && ld_dv >= 256,
                  "gradient pitches");
  const int nkv = ceil_div(N, BT);
  DESTR_CHECK_ARG(words_per_row >= nkv * 4, "words_per_row");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t qrows = static_cast<uint64_t>(B) * Q, krows = static_cast<uint64_t>(B) * N;
  CUtensorMap tqo, tqp, tke, tkp, tv, tdo;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tqo, q_obj, qrows, 512, 512, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tqp, q_pos, qrows, 256, 256, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tke, k_enc, krows, 256, ld_kenc, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tkp, k_pos, krows, 256, ld_kpos, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, v, krows, 256, ld_v, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, dout, qrows, 512, 512, BT, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const size_t smem = sizeof(SmemF) + 1024;
  DESTR_SMEM_OPTIN(cross_attn_bwd_fused_kernel, smem);
  cross_delta_kernel<<<ceil_div((int)qrows, 8), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out),
                                                              static_cast<const __nv_bfloat16*>(dout), delta, B, Q);
  DESTR_LAUNCH_CHECK();
  cross_attn_bwd_fused_kernel<<<dim3(nkv, B), NTHREADS_F, smem, st>>>(
      tqo, tqp, tke, tkp, tv, tdo, mask_bits, words_per_row, lse, delta, static_cast<__nv_bfloat16*>(dS_all),
      static_cast<__nv_bfloat16*>(dS_sum), static_cast<__nv_bfloat16*>(dk_enc), ld_dke,
      static_cast<__nv_bfloat16*>(dk_pos), ld_dkp, static_cast<__nv_bfloat16*>(dv), ld_dv, Q, N, nkv * BT, scale,
      scale * 1.4426950408889634f, Drop{drop_seed, drop_thr16, drop_site});
  DESTR_LAUNCH_CHECK();
  return 0;
}
