// Encoder multi-head self-attention backward (d_head = 32) on tcgen05 / TMEM / TMA.
//
// Backward of the fused attention in enc_attn_fwd.cu (reference arithmetic: torch
// F.multi_head_attention_forward as called at src/model/blocks/encoder_block.py:97-103).
// Scores are recomputed from Q,K and the saved log-sum-exp; nothing N x N is ever stored.
//
// Work item = one 128-key tile j of one (batch, head); a persistent CTA (one per SM, 608 threads) walks
// its items and, inside an item, the queries in 64-row sub-tiles u.  Everything is computed TRANSPOSED
// (rows = keys) so that thread <-> key row <-> TMEM lane:
//   S^T_u  = K_j . Q_u^T            (SS MMA, N = 64)      -> TMEM ST[u % 3]
//   dP^T_u = V_j . dO_u^T           (SS MMA, N = 64)      -> TMEM DPT[u % 3]
//   P^T = exp2(c*S^T - lse_q),  dS^T = P^T o (dP^T - delta_q)          (compute group g = u & 1)
//   dV_j += P^T_u . dO_u            (TS MMA, A = P^T bf16 in TMEM, in place of S^T; B = dO_u MN-major)
//   dK_j += dS^T_u . Q_u            (SS MMA, A = dS^T in smem K-major SW128, B = Q_u MN-major)
//   dQ_p  = dS_p . K_j              once per PAIR p of sub-tiles (M = 128 queries): A = the two dS^T blocks
//                                   of the pair read MN-major, B = K_j MN-major
// Two compute groups of 8 warps alternate on the sub-tiles (group g = u & 1).  S^T / dP^T live in THREE TMEM
// buffers (u % 3): the refill with sub-tile u+3 is issued right behind dV(u) by the same thread, i.e. a whole group
// iteration before the group that owns u+3 asks for it -- no compute warp waits for an
// arrive -> MMA warp -> tensor pipe -> commit round trip (with two buffers that wait was ~25 % of the warps'
// time).  P^T (bf16) overwrites the columns of S^T its own warp has already read (in-place, like the
// forward), which is what pays for the third buffer.  Warp w of a group: TMEM lanes 32*(w%4).., query
// columns 32*((w/4)%2).. of the sub-tile.  Two MMA-issuing warps (see there) share the ~20 small MMAs per sub-tile.
// dQ_p is pulled out of TMEM one pair later by group p&1: registers -> the warp's own smem tile -> one TMA reduce-add
// (fp32) per warp into dq_acc, which a small kernel converts to bf16; dV_j / dK_j are stored right after the first
// sub-tile of the next item (the next item's first gradient MMA waits for that drain: one bubble per
// 2*ceil(N/128) sub-tiles); the TMA / MMA warps run ahead across item boundaries.
// Measurements behind this shape: profiles/r02_enc_attn_bwd_knockout.txt (destr_debug_knob 18 = the knock-out bits).
//
// TMEM (512 columns): buffer b = u % 3 at 128 b: ST [0,64) DPT [64,128), P^T of the warp owning query columns
//                     32 cg.. in ST columns [32 cg, 32 cg + 16);  DV [384,416) DK [416,448) DQ0 [448,480) DQ1 [480,512)
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {
extern int g_knobs[24];
namespace {

constexpr int DH = 32;
constexpr int BT = 128;  // keys per item, queries per pair
constexpr int BQ = 64;   // queries per sub-tile
constexpr int QSTAGES = 8;
constexpr int NTHREADS = 608;  // 16 compute warps + TMA warp + 2 MMA-issuing warps
constexpr uint32_t KV_BYTES = BT * DH * 2;       // 8192
constexpr uint32_t Q_BYTES = BQ * DH * 2;        // 4096
constexpr uint32_t STAT_BYTES = 2 * BQ * 4;      // 512: -lse[64] | -delta[64]
constexpr uint32_t DS_BLOCK_BYTES = BT * 128;    // 16384: [128 key rows][64 queries] bf16, SW128
constexpr uint32_t C_ST = 0, C_DPT = 64, C_SBUF = 128, C_DV = 384, C_DK = 416, C_DQ = 448;

struct __align__(1024) Smem {
  uint8_t k[2][KV_BYTES];
  uint8_t v[2][KV_BYTES];
  uint8_t q[QSTAGES][Q_BYTES];
  uint8_t d_o[QSTAGES][Q_BYTES];
  uint8_t ds[2][2][DS_BLOCK_BYTES];  // [pair parity][sub-tile within the pair]
  float dq_stage[2][BT * DH];        // [group][warp]: 32 queries x 16 floats (SW64) on their way to the TMA reduce-add
  float stat[QSTAGES][2 * BQ];
  uint64_t kv_full[2], kv_free[2];
  uint64_t q_full[QSTAGES], q_empty[QSTAGES];
  uint64_t sdp_full[3], pds_full[3];  // S^T/dP^T ready, P^T/dS^T written: both by TMEM buffer (u % 3) -- a group can
                                      // never be two phases ahead of the MMA warp on a buffer (it can on its own parity)
  uint64_t dq_done[2], dq_free[2];
  uint64_t dkv_full, dkv_free;
  uint32_t tmem_base;
};

struct Knobs {
  uint32_t mn64_lbo, mn64_sbo, kmaj_lbo, a_mn_lbo, a_mn_sbo, a_mn_kstep;
  uint32_t dbg;  // timing experiments only (knob 18): 1 no dV, 2 no dK, 4 no dQ, 8 no softmax-backward math, 16 no S/dP, 32 no dQ atomics, 64 no dK/dV stores
};

struct Item {
  int b, h, j;
};
__device__ __forceinline__ Item decode_item(int w, int nkt, int heads) {
  Item it;
  it.j = w % nkt;
  const int bh = w / nkt;
  it.h = bh % heads;
  it.b = bh / heads;
  return it;
}

template <int POLYQ, bool DROP>
__global__ void __launch_bounds__(NTHREADS, 1)
enc_attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_dq, const uint32_t* __restrict__ mask_bits,
                    int words_per_row, const float* __restrict__ stats, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
                    int ld_dk, int ld_dv, int N, int heads, int n_items, float scale, float scale_log2, Knobs kn,
                    DropBits dp) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkt = (N + BT - 1) / BT;  // key tiles per (b,h) == query pairs per item
  const int npairs = nkt;
  const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int Ptot = my_items * npairs;  // flat pair count; flat sub-tile U = 2*Pf + g

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.kv_full[s], 1);
      mbar_init(&sm.kv_free[s], 2);
      mbar_init(&sm.dq_done[s], 1);
      mbar_init(&sm.dq_free[s], 8);
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(&sm.sdp_full[s], 1);
      mbar_init(&sm.pds_full[s], 8);
    }
    mbar_init(&sm.dkv_full, 2);
    mbar_init(&sm.dkv_free, 16);
    for (int s = 0; s < QSTAGES; ++s) {
      mbar_init(&sm.q_full[s], 1);
      mbar_init(&sm.q_empty[s], 2);
    }
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 16) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int U = 0;
      for (int it = 0; it < my_items; ++it) {
        const Item im = decode_item(blockIdx.x + it * gridDim.x, nkt, heads);
        const int row_base = im.b * N;
        const int kb = it & 1;
        mbar_wait_backoff(&sm.kv_free[kb], ((it >> 1) & 1) ^ 1, 1);
        mbar_arrive_expect_tx(&sm.kv_full[kb], 2 * KV_BYTES);
        tma_load_2d(sm.k[kb], &tm_k, &sm.kv_full[kb], im.h * DH, row_base + im.j * BT);
        tma_load_2d(sm.v[kb], &tm_v, &sm.kv_full[kb], im.h * DH, row_base + im.j * BT);
        const float* st_bh = stats + (static_cast<size_t>(im.b) * heads + im.h) * (2 * npairs) * (2 * BQ);
        for (int u = 0; u < 2 * npairs; ++u, ++U) {
          const int s = U % QSTAGES;
          mbar_wait(&sm.q_empty[s], ((U / QSTAGES) & 1) ^ 1, 2);
          if (kn.dbg & 128) {
            mbar_arrive_expect_tx(&sm.q_full[s], STAT_BYTES);
          } else {
            mbar_arrive_expect_tx(&sm.q_full[s], 2 * Q_BYTES + STAT_BYTES);
            tma_load_2d(sm.q[s], &tm_q, &sm.q_full[s], im.h * DH, row_base + u * BQ);
            tma_load_2d(sm.d_o[s], &tm_do, &sm.q_full[s], im.h * DH, row_base + u * BQ);
          }
          bulk_load(sm.stat[s], st_bh + static_cast<size_t>(u) * (2 * BQ), STAT_BYTES, &sm.q_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 17) {
    // ------------------------------ MMA issuers ------------------------------
    // At d_head = 32 the MMAs are small (16-64 tensor cycles each) and ~20 of them, 3 commits and 3-4 barrier waits
    // are needed per sub-tile: ONE issuing thread took ~1100 cycles per sub-tile for that (measured with clock64),
    // co-critical with the compute groups.  So two single-thread issuers on different SM sub-partitions:
    //   warp 17  dV(U) then S^T / dP^T of sub-tile U+3 into the TMEM buffer dV(U) has just read P^T from -- same
    //            thread, so the tensor pipe orders them and the refill needs no barrier round trip
    //   warp 18  dK(U) and, every second sub-tile, dQ of the pair
    // MMAs of different issuers are only ordered through mbarriers (tcgen05.commit), never through program order.
    if (elect_one()) {
      constexpr uint32_t id_sT = umma_idesc_bf16(BT, BQ, false, false);   // K-major x K-major, N = 64
      constexpr uint32_t id_kn = umma_idesc_bf16(BT, DH, false, true);    // A K-major (TMEM/smem), B MN-major
      constexpr uint32_t id_nn = umma_idesc_bf16(BT, DH, true, true);     // A MN-major, B MN-major
      constexpr uint64_t D_KMAJ64 = umma_desc_const(16, 512, SWZ_64B);      // K-major, 64-byte rows
      constexpr uint64_t D_MN64 = umma_desc_const(512, 512, SWZ_64B);       // MN-major, 64-byte rows (k-step 1024 B)
      constexpr uint64_t D_KMAJ128 = umma_desc_const(16, 1024, SWZ_128B);   // dS^T as K-major A (k-step 32 B)
      constexpr uint64_t D_AMN128 = umma_desc_const(16384, 1024, SWZ_128B); // dS^T pair as MN-major A (k-step 2048 B)
      const uint32_t a_k = smem_u32(sm.k[0]) >> 4, a_v = smem_u32(sm.v[0]) >> 4, a_q = smem_u32(sm.q[0]) >> 4,
                     a_do = smem_u32(sm.d_o[0]) >> 4, a_ds = smem_u32(sm.ds[0][0]) >> 4;
      const int Utot = 2 * Ptot, SUB = 2 * npairs;
      if (warp == 17) {
        // look-ahead cursor of the S^T / dP^T issue: flat sub-tile Uq = sub-tile uq of item itq, TMEM buffer bq
        int Uq = 0, itq = 0, uq = 0, bq = 0;
        auto issue_sdp = [&]() {
          const uint32_t s = Uq % QSTAGES, kb = itq & 1;
          if (uq == 0) mbar_wait_backoff(&sm.kv_full[kb], (itq >> 1) & 1, 3);
          mbar_wait_backoff(&sm.q_full[s], (Uq / QSTAGES) & 1, 4);
          tc_fence_after();
          const uint64_t dk_ = D_KMAJ64 + (a_k + kb * (KV_BYTES >> 4)), dv_ = D_KMAJ64 + (a_v + kb * (KV_BYTES >> 4));
          const uint64_t dq_ = D_KMAJ64 + (a_q + s * (Q_BYTES >> 4)), ddo_ = D_KMAJ64 + (a_do + s * (Q_BYTES >> 4));
          const uint32_t t = tmem + bq * C_SBUF;
          if (!(kn.dbg & 16)) {
            umma_ss(t + C_ST, dk_, dq_, id_sT, 0u);
            umma_ss(t + C_ST, dk_ + 2, dq_ + 2, id_sT, 1u);
            umma_ss(t + C_DPT, dv_, ddo_, id_sT, 0u);
            umma_ss(t + C_DPT, dv_ + 2, ddo_ + 2, id_sT, 1u);
          }
          tc_commit(&sm.sdp_full[bq]);
          ++Uq;
          if (++bq == 3) bq = 0;
          if (++uq == SUB) {
            tc_commit(&sm.kv_free[kb]);  // (1 of 2 arrivals) every S^T / dP^T of the item has been issued
            uq = 0;
            ++itq;
          }
        };
        for (int i = 0; i < 3 && Uq < Utot; ++i) issue_sdp();
        int it = 0, u = 0, b3 = 0, ph3 = 0;
        for (int U = 0; U < Utot; ++U) {
          const uint32_t s = U % QSTAGES;
          mbar_wait(&sm.pds_full[b3], ph3, 5);  // the group has written P^T (TMEM, in place of S^T) and dS^T (smem)
          if (u == 0 && it > 0) mbar_wait_backoff(&sm.dkv_free, (it - 1) & 1, 6);  // dV / dK of item it-1 drained
          tc_fence_after();
          const uint64_t ddo_mn = D_MN64 + (a_do + s * (Q_BYTES >> 4));
          const uint32_t t_pt = tmem + b3 * C_SBUF + C_ST;
          if (!(kn.dbg & 1))
#pragma unroll
            for (int ks = 0; ks < BQ / 16; ++ks)  // dV += P^T_u dO_u; k-step ks = 16 queries: warp column cg = ks / 2, chunk ks % 2
              umma_ts(tmem + C_DV, t_pt + (ks >> 1) * 32 + (ks & 1) * 8, ddo_mn + ks * 64, id_kn,
                      (ks > 0 || u > 0) ? 1u : 0u);
          tc_commit(&sm.q_empty[s]);  // (1 of 2)
          if (++u == SUB) {
            tc_commit(&sm.dkv_full);  // (1 of 2)
            u = 0;
            ++it;
          }
          if (Uq < Utot) issue_sdp();  // refill the buffer dV(U) has just read: ordered behind it in the tensor pipe
          if (++b3 == 3) { b3 = 0; ph3 ^= 1; }
        }
      } else if (warp == 18) {
        int it = 0, u = 0, b3 = 0, ph3 = 0;
        for (int U = 0; U < Utot; ++U) {
          const uint32_t g = U & 1, s = U % QSTAGES, Pf = U >> 1, pb = Pf & 1, kb = it & 1;
          mbar_wait(&sm.pds_full[b3], ph3, 5);
          if (u == 0 && it > 0) mbar_wait_backoff(&sm.dkv_free, (it - 1) & 1, 6);
          tc_fence_after();
          const uint64_t dq_mn = D_MN64 + (a_q + s * (Q_BYTES >> 4));
          const uint64_t dds_k = D_KMAJ128 + (a_ds + (pb * 2 + g) * (DS_BLOCK_BYTES >> 4));
          if (!(kn.dbg & 2))
#pragma unroll
            for (int ks = 0; ks < BQ / 16; ++ks)  // dK += dS^T_u Q_u
              umma_ss(tmem + C_DK, dds_k + ks * 2, dq_mn + ks * 64, id_kn, (ks > 0 || u > 0) ? 1u : 0u);
          tc_commit(&sm.q_empty[s]);  // (2 of 2)
          if (g == 1) {
            // dQ_pair = dS_pair K_j   (M = 128 queries spanning the pair's two dS^T blocks)
            mbar_wait_backoff(&sm.dq_free[pb], ((Pf >> 1) & 1) ^ 1, 7);
            tc_fence_after();
            const uint64_t dds_mn = D_AMN128 + (a_ds + pb * 2 * (DS_BLOCK_BYTES >> 4));
            const uint64_t dk_mn = D_MN64 + (a_k + kb * (KV_BYTES >> 4));
            if (!(kn.dbg & 4))
#pragma unroll
              for (int ks = 0; ks < BT / 16; ++ks)
                umma_ss(tmem + C_DQ + pb * 32, dds_mn + ks * 128, dk_mn + ks * 64, id_nn, ks > 0);
            tc_commit(&sm.dq_done[pb]);  // dQ and both dK of the pair have completed: the pair's dS^T blocks are free
          }
          if (++u == SUB) {
            tc_commit(&sm.dkv_full);     // (2 of 2)
            tc_commit(&sm.kv_free[kb]);  // (2 of 2) K_j is no longer needed
            u = 0;
            ++it;
          }
          if (++b3 == 3) { b3 = 0; ph3 ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ compute warps ------------------------------
    const int g = warp >> 3;              // compute group: sub-tile parity
    const int wq = warp & 3, cg = (warp >> 2) & 1;
    const int krow = wq * 32 + lane;      // TMEM lane: key row (S^T, dV, dK) / query row (dQ)
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2);

    // dQ of flat pair Df (= pair `pair` of item im; parity == g): TMEM -> registers -> the warp's own smem tile (32
    // queries x 16 floats, SW64) -> one TMA reduce-add (fp32) per warp into dq_acc; nothing but the warp itself is
    // involved, so no block-level barrier.  (Per-thread red.global.add: every lane hits a different 128-byte line,
    // 4.8 M 16-byte atomics per launch, and the warp stalls on their operand registers at its next iteration.)
    // Padding query rows of the last pair (>= N) carry exact zeros (P = 0 there), so adding them into the next
    // image's rows is harmless; rows past the end of the tensor are clipped by the tensor map.
    const uint32_t stage_addr = smem_u32(sm.dq_stage[g]) + (wq * 2 + cg) * 2048;
    auto drain_dq = [&](int Df, const Item& im, int pair) {
      const int pb = Df & 1;
      mbar_wait(&sm.dq_done[pb], (Df >> 1) & 1, 8);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_x16(tmem + lane_addr + C_DQ + pb * 32 + cg * 16, r);
      tc_wait_ld();
      tc_fence_before();
      if (lane == 0) bulk_wait_group_read0();  // the warp's previous reduce has read the tile
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.dq_free[pb]);
      const uint32_t row_addr = stage_addr + lane * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts_u4(row_addr + ((c ^ ((lane >> 1) & 3)) << 4), __float_as_uint(__uint_as_float(r[4 * c]) * scale),
               __float_as_uint(__uint_as_float(r[4 * c + 1]) * scale), __float_as_uint(__uint_as_float(r[4 * c + 2]) * scale),
               __float_as_uint(__uint_as_float(r[4 * c + 3]) * scale));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && !(kn.dbg & 32)) {
        tma_reduce_add_2d(&tm_dq, stage_addr, im.h * DH + cg * 16, im.b * N + pair * BT + wq * 32);
        bulk_commit_group();
      }
    };
    auto drain_dkv = [&](int it, const Item& im) {  // group 0 stores dV_j, group 1 stores dK_j
      mbar_wait(&sm.dkv_full, it & 1, 9);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_x16(tmem + lane_addr + (g == 0 ? C_DV : C_DK) + cg * 16, r);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.dkv_free);
      const int key = im.j * BT + krow;
      if (key < N && !(kn.dbg & 64)) {
        // dV = (dropped P)^T dO: the 1/(1-p) of the forward's dropout is applied here, once per output
        const float f = (g == 0) ? (DROP ? drop_scale(dp.thr16) : 1.f) : scale;
        uint32_t o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = pack_bf16x2(__uint_as_float(r[2 * c]) * f, __uint_as_float(r[2 * c + 1]) * f);
        __nv_bfloat16* base = (g == 0) ? dv + static_cast<size_t>(im.b * N + key) * ld_dv
                                       : dk + static_cast<size_t>(im.b * N + key) * ld_dk;
        uint4* dst = reinterpret_cast<uint4*>(base + im.h * DH + cg * 16);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    };

    const uint32_t sb = smem_u32(&sm);
    const uint32_t b_sdp = sb + offsetof(Smem, sdp_full), b_pds = sb + offsetof(Smem, pds_full);
    const uint32_t b_qfull = sb + offsetof(Smem, q_full), b_dqdone = sb + offsetof(Smem, dq_done);
    const uint32_t a_stat = sb + offsetof(Smem, stat) + cg * 128;                 // -lse of my 32 queries (+256: -delta)
    const uint32_t a_ds = sb + offsetof(Smem, ds) + g * DS_BLOCK_BYTES + krow * 128;  // my dS^T row (pair parity 0)
    const uint32_t t_mine = tmem + lane_addr + C_ST + cg * 32;  // my 32 query columns of S^T (buffer 0)
    const uint32_t swz = krow & 7;
    const float drop_s = drop_scale(dp.thr16);
    const uint64_t ds2s = pack_f32x2(drop_s, drop_s);

    Item prev{0, 0, 0};
    uint32_t Pf = 0;
    uint32_t b3 = g, ph3 = 0;  // TMEM buffer (U % 3) and barrier parity ((U / 3) & 1) of my sub-tile U = 2 Pf + g
    for (int it = 0; it < my_items; ++it) {
      const Item im = decode_item(blockIdx.x + it * gridDim.x, nkt, heads);
      const int key = im.j * BT + krow;
      const bool key_masked = (mask_bits[static_cast<size_t>(im.b) * words_per_row + (key >> 5)] >> (key & 31)) & 1u;
      // dropout bits of my key row: word w <-> queries [32 w, 32 w + 32) = sub-tile w / 2, half cg
      const uint32_t* dw = DROP ? dp.bits + (static_cast<size_t>(im.b * heads + im.h) * dp.words + g * 2 + cg) * (nkt * BT) + key
                                : mask_bits;
      uint32_t dnext = DROP ? dw[0] : 0u;
      for (int p = 0; p < npairs; ++p, ++Pf) {
        const uint32_t dbits = dnext;
        if (DROP && p + 1 < npairs) dnext = dw[static_cast<size_t>(4 * (p + 1)) * (nkt * BT)];
        const uint32_t U = 2 * Pf + g, s = U % QSTAGES, pb = Pf & 1;
        mbar_wait_a(b_sdp + b3 * 8, ph3, 10);
        mbar_wait_a(b_qfull + s * 8, (U / QSTAGES) & 1, 11);  // long complete: makes the TMA-written stats visible
        tc_fence_after();
        const uint32_t t_st = t_mine + b3 * C_SBUF, b_pds_u = b_pds + b3 * 8;
        b3 += 2;
        if (b3 >= 3) { b3 -= 3; ph3 ^= 1; }
        const uint32_t stat_s = a_stat + s * STAT_BYTES;
        const uint32_t ds_row = a_ds + pb * (2 * DS_BLOCK_BYTES);
        if (kn.dbg & 8) mbar_wait_a(b_dqdone + pb * 8, ((Pf >> 1) & 1) ^ 1, 12);
        if (!(kn.dbg & 8))
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {  // two chunks of 16 query columns (keeps the live set small)
          uint32_t st[16], dpv[16];
          tmem_ld_x16(t_st + hc * 16, st);
          tmem_ld_x16(t_st + (C_DPT - C_ST) + hc * 16, dpv);
          tc_wait_ld();
          if (hc == 0) mbar_wait_a(b_dqdone + pb * 8, ((Pf >> 1) & 1) ^ 1, 12);  // dS^T block free: dQ(Pf-2) has read it
          uint32_t pk[8], dsk[8];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 a = lds_f4(stat_s + (hc * 4 + c4) * 16), d = lds_f4(stat_s + BQ * 4 + (hc * 4 + c4) * 16);
            const float nl[4] = {a.x, a.y, a.z, a.w}, nd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int c = c4 * 4 + e;
              const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(st[c]), __uint_as_float(st[c + 1])), sc2,
                                            pack_f32x2(nl[e], nl[e + 1]));
              float x0, x1, p0, p1;
              unpack_f32x2(x2, x0, x1);
              if (((c >> 1) & 3) < POLYQ) {
                ex2_poly_f32x2(x0, x1, p0, p1);
              } else {
                p0 = ex2_approx(x0);
                p1 = ex2_approx(x1);
              }
              float g0 = __uint_as_float(dpv[c]), g1 = __uint_as_float(dpv[c + 1]), pd0 = p0, pd1 = p1;
              if (DROP) {  // the forward dropped P (and rescaled): dP and the P that multiplies dO go through the mask
                if ((dbits >> (hc * 16 + c)) & 1u) g0 = pd0 = 0.f;      // bit = query column of this thread
                if ((dbits >> (hc * 16 + c + 1)) & 1u) g1 = pd1 = 0.f;  // (the 1/(1-p) of P^T is applied to dV at the drain)
              }
              const uint64_t dd = DROP ? fma_f32x2(pack_f32x2(g0, g1), ds2s, pack_f32x2(nd[e], nd[e + 1]))
                                       : add_f32x2(pack_f32x2(g0, g1), pack_f32x2(nd[e], nd[e + 1]));
              const uint64_t ds2 = mul_f32x2(pack_f32x2(p0, p1), dd);
              float d0, d1;
              unpack_f32x2(ds2, d0, d1);
              pk[c >> 1] = pack_bf16x2(pd0, pd1);
              dsk[c >> 1] = pack_bf16x2(d0, d1);
            }
          }
          if (key_masked) {  // a padded key contributes nothing: P^T row = dS^T row = 0
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = dsk[i] = 0u;
          }
          tmem_st_x8(t_st + hc * 8, pk);  // P^T over S^T columns this warp has already read
#pragma unroll
          for (int c = 0; c < 2; ++c)  // 8 queries per 16-byte chunk; chunk index = 4*cg + 2*hc + c
            sts_u4(ds_row + (((4 * cg + 2 * hc + c) ^ swz) << 4), dsk[4 * c], dsk[4 * c + 1], dsk[4 * c + 2],
                   dsk[4 * c + 3]);
        }
        if (!(kn.dbg & 2048)) tc_wait_st();
        if (!(kn.dbg & 1024)) fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(b_pds_u);  // one arrival per warp

        // ---- deferred drains (their MMAs were issued one pair / one item ago) ----
        if (Pf > 0 && ((Pf - 1) & 1) == static_cast<uint32_t>(g)) {
          if (p > 0) drain_dq(Pf - 1, im, p - 1);
          else drain_dq(Pf - 1, prev, npairs - 1);
        }
        if (p == 0 && it > 0) drain_dkv(it - 1, prev);
      }
      prev = im;
    }
    if (Ptot > 0) {
      if (((Ptot - 1) & 1) == g) drain_dq(Ptot - 1, prev, npairs - 1);
      drain_dkv(my_items - 1, prev);
    }
    if (lane == 0) bulk_wait_group0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tmem);
}

// Pre-pass, one warp per (padded) token row q of an image, Np = 128*ceil(N/128) rows per image:
//   delta[b,h,q] = sum_d dO[q,h,d] * O[q,h,d];  stats[b,h][q/64][0][q%64] = -lse, [1][q%64] = -delta
//   (padding rows: -lse = -inf -> P = 0, -delta = 0);  dq_acc row q <- 0.
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                     const float* __restrict__ lse, float* __restrict__ stats,
                                     float* __restrict__ dq_acc, int B, int N, int Np, int heads) {
  const int lane = threadIdx.x & 31;
  const int prow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (prow >= B * Np) return;
  const int b = prow / Np, q = prow - b * Np;
  const int cols = heads * DH;
  const int hh = lane >> 2;
  float s = 0.f;
  if (q < N && lane * 8 < cols) {
    const size_t row = static_cast<size_t>(b) * N + q;
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * cols + lane * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(d_o + row * cols + lane * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(gw[i] << 16), s);
      s = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(gw[i] & 0xffff0000u), s);
    }
    float4* z = reinterpret_cast<float4*>(dq_acc + row * cols + lane * 8);
    z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if ((lane & 3) == 0 && hh < heads) {
    float* dst = stats + ((static_cast<size_t>(b) * heads + hh) * (Np / BQ) + q / BQ) * (2 * BQ) + (q % BQ);
    dst[0] = (q < N) ? -lse[(static_cast<size_t>(b) * heads + hh) * N + q] : -INFINITY;
    dst[BQ] = (q < N) ? -s : 0.f;
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

extern "C" int destr_enc_attn_bwd_stats_floats(int B, int N, int heads) {
  return B * heads * ((N + 127) / 128) * 128 * 2;
}

extern "C" int destr_enc_attn_bwd(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v,
                                  const uint32_t* mask_bits, int words_per_row, const void* out, const void* dout,
                                  const float* lse, float* stats, float* dq_acc, void* dq, void* dk, void* dv,
                                  int ld_dq, int ld_dk, int ld_dv, int B, int N, int heads, float scale,
                                  const uint32_t* drop_colbits, int drop_words, uint32_t drop_thr16, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(q && k && v && mask_bits && out && dout && lse && stats && dq_acc && dq && dk && dv, "null pointer");
  DESTR_CHECK_ARG(B > 0 && N > 0 && heads > 0 && heads * DH <= 256, "shape");
  DESTR_CHECK_ARG(words_per_row >= ceil_div(N, BT) * 4, "words_per_row");
  DESTR_CHECK_ARG(!drop_thr16 || (drop_colbits && drop_words >= ceil_div(N, BT) * 4),
                  "dropout needs the column bit matrix of destr_attn_dropout_bits");
  DESTR_CHECK_ARG(ld_dq % 8 == 0 && ld_dk % 8 == 0 && ld_dv % 8 == 0, "gradient row pitch must be a multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t rows = static_cast<uint64_t>(B) * N;
  const int cols = heads * DH;
  const int Np = ceil_div(N, BT) * BT;
  CUtensorMap tq, tk, tv, tdo, tdq;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, q, rows, cols, ld_q, BQ, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, k, rows, cols, ld_k, BT, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, v, rows, cols, ld_v, BT, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, dout, rows, cols, cols, BQ, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_f32_2d(&tdq, dq_acc, rows, cols, cols, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  const size_t smem = sizeof(Smem) + 1024;
  using KernelT = decltype(&enc_attn_bwd_kernel<0, false>);
  static const KernelT kernels[8] = {enc_attn_bwd_kernel<0, false>, enc_attn_bwd_kernel<1, false>,
                                     enc_attn_bwd_kernel<2, false>, enc_attn_bwd_kernel<3, false>,
                                     enc_attn_bwd_kernel<0, true>,  enc_attn_bwd_kernel<1, true>,
                                     enc_attn_bwd_kernel<2, true>,  enc_attn_bwd_kernel<3, true>};
  for (int i = 0; i < 8; ++i) DESTR_SMEM_OPTIN(kernels[i], smem);
  const bool main_only = g_knobs[14] != 0;  // bench: time the tcgen05 kernel alone (outputs are then meaningless)
  if (!main_only)
    attn_bwd_prep_kernel<<<ceil_div(B * Np, 8), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out),
                                                            static_cast<const __nv_bfloat16*>(dout), lse, stats,
                                                            dq_acc, B, N, Np, heads);
  DESTR_LAUNCH_CHECK();
  Knobs kn{(uint32_t)g_knobs[0], (uint32_t)g_knobs[1], (uint32_t)g_knobs[2],
           (uint32_t)g_knobs[6], (uint32_t)g_knobs[7], (uint32_t)g_knobs[8], (uint32_t)g_knobs[18]};
  const int n_items = B * heads * ceil_div(N, BT);
  int grid = n_items < 148 ? n_items : 148;  // persistent: one CTA per SM
  if (g_knobs[13] > 0 && g_knobs[13] < grid) grid = g_knobs[13];
  kernels[(g_knobs[10] & 3) + (drop_thr16 ? 4 : 0)]<<<grid, NTHREADS, smem, st>>>(
      tq, tk, tv, tdo, tdq, mask_bits, words_per_row, stats, static_cast<__nv_bfloat16*>(dk),
      static_cast<__nv_bfloat16*>(dv), ld_dk, ld_dv, N, heads, n_items, scale, scale * 1.4426950408889634f, kn,
      DropBits{drop_colbits, drop_thr16, drop_words});
  DESTR_LAUNCH_CHECK();
  const int64_t n4 = static_cast<int64_t>(rows) * cols / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (!main_only)
    cvt_f32_bf16_rows_kernel<<<blocks, 256, 0, st>>>(dq_acc, static_cast<__nv_bfloat16*>(dq), (int)rows, cols, ld_dq);
  DESTR_LAUNCH_CHECK();
  return 0;
}
