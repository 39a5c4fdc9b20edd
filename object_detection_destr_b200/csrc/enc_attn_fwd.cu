// Encoder multi-head self-attention forward (d_head = 32), flash style, on tcgen05 / TMEM / TMA.
//
// Replaces the bmm -> softmax -> bmm chain of torch's F.multi_head_attention_forward as invoked by
// the reference at src/model/blocks/encoder_block.py:97-103 (need_weights=True math path: q*scale,
// QK^T, key-padding mask -> -inf, softmax, PV) without ever materialising the N x N scores.
//
// At d_head = 32 the exp (MUFU) pipe and the issue slots of the softmax warps bound this kernel, not
// the tensor pipe (a 128x96 score tile = 12 k exps = 768 MUFU cycles/SM against 192 tensor cycles), so
// everything is arranged to keep four softmax warps per SM sub-partition busy:
//   * persistent CTAs, 2 per SM, 320 threads: work item = 128 query rows of one (batch, head); the TMA
//     and MMA warps run ahead across item boundaries, so an item's prologue hides under the previous
//     item's softmax.
//   * warps 0-7 softmax: thread <-> query row (TMEM lane 32*(warp%4)+lane); warps 0-3 own keys [0,48) of
//     every 96-key tile, warps 4-7 keys [48,96).  Each half keeps its own running max / sum and its own
//     O accumulator (split-K); the halves merge once per item through shared memory.
//   * O stays in TMEM: P.V accumulates in place over the key tiles and is rescaled only when a row's
//     running max grows by more than 2^8 (lazy rescale) -- no per-tile O read-back, no accumulator
//     registers, which is what lets 640 threads/SM fit the register file.
//   * S is double buffered (Q.K_{j+1}^T runs under softmax(S_j)); packed fp32x2 FMA/ADD, 3-input max;
//     a knob moves a fraction of the exps from MUFU to an FMA-pipe polynomial (ex2_poly_f32x2).
//   * the softmax reads its 48 scores twice, 16 at a time (max, then exp): TMEM reads are cheap and the live set stays
//     small enough for 96 registers (holding all 48 made the dropout variant spill around its mask-word prefetch).
//   warp 8 TMA producer (Q double buffered; K_j / V_j through a KSTAGES ring), warp 9 tcgen05.mma issuer.
// TMEM (256 columns per CTA): S0 [0,96) S1 [96,192) O_a [192,224) O_b [224,256).  P (bf16) overwrites
// the first 24 columns of each half's own S region and is the A operand of P.V straight from TMEM.
// Head split is done by TMA: Q/K/V tiles are rows x 32 boxes of the token-major projection output
// (row pitch ld_*), 64-byte rows, SWIZZLE_64B.  Q,K are K-major UMMA operands; V is the MN-major B
// operand of P.V.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {

// debug / tuning knobs (destr_debug_knob): 0 v_lbo 1 v_sbo 2 qk_lbo 3 qk_sbo 4 p_kstep_cols 5 v_kstep_bytes
// 6-8 enc_attn_bwd descriptors, 9 enc fwd polynomial-exp quarter count (0..3), 10 enc bwd ditto,
// 11 lazy-rescale tau+1, 12/13 grid overrides, 14 enc bwd: launch the main kernel only (bench timing)
// 15 gemm_dw split override, 16 programmatic dependent launch: -1 = DESTR_PDL from the environment, 0 off, 1 on
// 17 gemm tile width: 1 / 2 force 128 / 256, 18 enc bwd knock-out bits, 19 enc fwd knock-out bits (timing experiments:
// tools/dbg_enc_attn_bwd.py, tools/dbg_enc_attn_fwd.py; results are meaningless with a bit set)
int g_knobs[24] = {512, 512, 16, 512, 8, 1024, 16384, 1024, 2048, 1, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0, 0};

namespace {

constexpr int DH = 32;
constexpr int BM = 128;
constexpr int BN = 96;
constexpr int HN = BN / 2;  // keys per softmax half
constexpr int KSTAGES = 4;
constexpr int NTHREADS = 320;
constexpr uint32_t Q_BYTES = BM * DH * 2;   // 8192
constexpr uint32_t KV_BYTES = BN * DH * 2;  // 6144
constexpr uint32_t C_S = 0, C_O = 2 * BN;   // TMEM columns
constexpr float kLazyTau = 8.f;             // rescale O only when the row max grows by > 2^8
constexpr int XCH = 20;                     // floats per row of the half-merge exchange (m, l, 16 x O, pad)

struct __align__(1024) Smem {
  uint8_t q[2][Q_BYTES];
  uint8_t k[KSTAGES][KV_BYTES];
  uint8_t v[KSTAGES][KV_BYTES];
  float xch[2][BM][XCH];
  uint64_t q_full[2];
  uint64_t q_empty[2];
  uint64_t kv_full[KSTAGES];
  uint64_t kv_empty[KSTAGES];
  uint64_t s_full[2];     // indexed by flat tile parity f & 1
  uint64_t p_full[2][2];  // [half][f & 1]
  uint64_t o_full[2];     // [half]: one phase per tile
  uint64_t o_free[2];     // [half]: one phase per item (epilogue has read O)
  uint64_t o_done[2];     // [half]: one phase per item (last P.V of the item has landed)
  uint32_t tmem_base;
};

struct Knobs {
  uint32_t v_lbo, v_sbo, qk_lbo, qk_sbo, p_kstep_cols, v_kstep_bytes;
  uint32_t dbg;  // timing experiments only (knob 19): 1 no Q.K^T MMAs, 2 no P.V MMAs, 4 no softmax math
};

// POLYQ of every 4 key pairs take the FMA-pipe polynomial exp instead of MUFU.EX2
template <int POLYQ, bool DROP>
__global__ void __launch_bounds__(NTHREADS, 2)
enc_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const uint32_t* __restrict__ mask_bits,
                    int words_per_row, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int heads,
                    int n_items, float scale_log2, float lazy_tau, Knobs kn, DropBits dp) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkv = (N + BN - 1) / BN;
  const int nqt = (N + BM - 1) / BM;
  // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...;  item -> ((b*heads + h)*nqt + qt)
  const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int T = my_items * nkv;  // flat tile count

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.q_full[s], 1);
      mbar_init(&sm.q_empty[s], 1);
      mbar_init(&sm.s_full[s], 1);
      mbar_init(&sm.p_full[0][s], BM);
      mbar_init(&sm.p_full[1][s], BM);
      mbar_init(&sm.o_full[s], 1);
      mbar_init(&sm.o_free[s], BM);
      mbar_init(&sm.o_done[s], 1);
    }
    for (int s = 0; s < KSTAGES; ++s) {
      mbar_init(&sm.kv_full[s], 1);
      mbar_init(&sm.kv_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<256>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 8) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int f = 0;
      for (int it = 0; it < my_items; ++it) {
        const int w = blockIdx.x + it * gridDim.x;
        const int qt = w % nqt, bh = w / nqt;
        const int h = bh % heads, b = bh / heads;
        const int row_base = b * N;
        const int qb = it & 1;
        mbar_wait_backoff(&sm.q_empty[qb], ((it >> 1) & 1) ^ 1, 1);
        mbar_arrive_expect_tx(&sm.q_full[qb], Q_BYTES);
        tma_load_2d(sm.q[qb], &tm_q, &sm.q_full[qb], h * DH, row_base + qt * BM);
        for (int j = 0; j < nkv; ++j, ++f) {
          const int s = f % KSTAGES;
          mbar_wait_backoff(&sm.kv_empty[s], ((f / KSTAGES) & 1) ^ 1, 2);
          mbar_arrive_expect_tx(&sm.kv_full[s], 2 * KV_BYTES);
          tma_load_2d(sm.k[s], &tm_k, &sm.kv_full[s], h * DH, row_base + j * BN);
          tma_load_2d(sm.v[s], &tm_v, &sm.kv_full[s], h * DH, row_base + j * BN);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------ MMA issuer ------------------------------
    // One thread issues every MMA and shares its scheduler with four softmax warps, so its instruction stream is
    // kept minimal: descriptors are compile-time constants plus a shifted address, item/tile counters are
    // incremental, and the operands of the look-ahead Q.K^T are awaited before the wait on P.
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BM, BN, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BM, DH, false, true);
      constexpr uint64_t D_KMAJ = umma_desc_const(16, 512, SWZ_64B);   // Q, K: K-major, 64-byte rows (k-step 32 B)
      constexpr uint64_t D_VMN = umma_desc_const(512, 512, SWZ_64B);   // V: MN-major B operand (k-step 16 keys = 1024 B)
      const uint32_t a_q = smem_u32(sm.q[0]) >> 4, a_k = smem_u32(sm.k[0]) >> 4, a_v = smem_u32(sm.v[0]) >> 4;
      // look-ahead cursor (item / tile of flat tile fq = the next Q.K^T to issue)
      int fq = 0, itq = 0, jq = 0;
      auto issue_qk = [&]() {  // S_{fq&1} = Q_item . K_j^T
        const uint32_t stage = fq % KSTAGES, qb = itq & 1;
        if (jq == 0) mbar_wait_backoff(&sm.q_full[qb], (itq >> 1) & 1, 3);
        mbar_wait_backoff(&sm.kv_full[stage], (fq / KSTAGES) & 1, 4);
        tc_fence_after();
        const uint64_t a = D_KMAJ + (a_q + qb * (Q_BYTES >> 4)), bd = D_KMAJ + (a_k + stage * (KV_BYTES >> 4));
        if (!(kn.dbg & 1)) {
          umma_ss(tmem + C_S + (fq & 1) * BN, a, bd, idesc_qk, 0u);
          umma_ss(tmem + C_S + (fq & 1) * BN, a + 2, bd + 2, idesc_qk, 1u);
        }
        tc_commit(&sm.s_full[fq & 1]);
        if (jq == nkv - 1) tc_commit(&sm.q_empty[qb]);  // every Q.K^T of the item has been issued
        ++fq;
        if (++jq == nkv) { jq = 0; ++itq; }
      };
      if (T > 0) issue_qk();
      if (T > 1) issue_qk();
      int it = 0, j = 0;
      for (int f = 0; f < T; ++f) {
        const uint32_t stage = f % KSTAGES;
        const uint64_t dv = D_VMN + (a_v + stage * (KV_BYTES >> 4));
        const uint32_t t_p = tmem + C_S + (f & 1) * BN;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {  // O_hf (+)= P_hf . V[keys of the half]
          mbar_wait(&sm.p_full[hf][f & 1], (f >> 1) & 1, 5);
          if (j == 0 && it > 0) mbar_wait_backoff(&sm.o_free[hf], (it - 1) & 1, 6);  // epilogue has read O
          tc_fence_after();
          if (!(kn.dbg & 2))
#pragma unroll
          for (int ks = 0; ks < HN / 16; ++ks)
            umma_ts(tmem + C_O + hf * DH, t_p + hf * HN + ks * 8, dv + (hf * (HN / 16) + ks) * 64, idesc_pv,
                    (j > 0 || ks > 0) ? 1u : 0u);
          tc_commit(&sm.o_full[hf]);
          if (j == nkv - 1) tc_commit(&sm.o_done[hf]);
        }
        tc_commit(&sm.kv_empty[stage]);
        if (fq < T) issue_qk();  // reuses S_{f&1}: ordered behind P.V(f) in the tensor pipe
        if (++j == nkv) { j = 0; ++it; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ softmax warps ------------------------------
    const int wq = warp & 3, hf = warp >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2);
    const int Np_drop = (N + 127) / 128 * 128;
    int f = 0;
    for (int it = 0; it < my_items; ++it) {
      const int w = blockIdx.x + it * gridDim.x;
      const int qt = w % nqt, bh = w / nqt;
      const int h = bh % heads, b = bh / heads;
      const uint32_t* mrow = mask_bits + static_cast<size_t>(b) * words_per_row;
      float m_ref = -INFINITY, l = 0.f;
      // dropout bits of my query row: same word / half split as the key mask (bit i <-> key 96 j + 48 hf + i)
      const uint32_t* dw = DROP ? dp.bits + (static_cast<size_t>(bh) * dp.words + hf) * Np_drop + qt * BM + row : mrow;
      uint32_t dw0 = 0, dw1 = 0;
      if (DROP) {
        dw0 = dw[0];
        dw1 = dw[Np_drop];
      }
      // mask bits of this half's 48 keys of tile j (bit i <-> key 96 j + 48 hf + i) live in two words; the raw
      // words are prefetched one tile ahead and only combined when consumed (no stall on the load)
      const uint32_t* mw = mrow + hf;
      uint32_t mw0 = mw[0], mw1 = mw[1];

      for (int j = 0; j < nkv; ++j, ++f) {
        const uint32_t s_addr = tmem + lane_addr + C_S + (f & 1) * BN + hf * HN;
        const uint32_t c0 = mw0, c1 = mw1;
        const uint32_t e0 = dw0, e1 = dw1;
        if (j + 1 < nkv) {
          mw0 = mw[3 * (j + 1)];
          mw1 = mw[3 * (j + 1) + 1];
          if (DROP) {
            dw0 = dw[static_cast<size_t>(3 * (j + 1)) * Np_drop];
            dw1 = dw[static_cast<size_t>(3 * (j + 1) + 1) * Np_drop];
          }
        }
        mbar_wait(&sm.s_full[f & 1], (f >> 1) & 1, 7);
        tc_fence_after();
        uint64_t rs2[2] = {0ull, 0ull};
        if (!(kn.dbg & 4)) {
        // hf 0: keys [0,48) = word0 | low half of word1;  hf 1: keys [48,96) = high half of word1 | word2
        const uint64_t mb = hf == 0 ? (static_cast<uint64_t>(c0) | (static_cast<uint64_t>(c1 & 0xffffu) << 32))
                                    : (static_cast<uint64_t>(c0 >> 16) | (static_cast<uint64_t>(c1) << 16));
        // Two passes over the half's 48 scores, 16 at a time (TMEM reads are cheap -- ~450 B/clk/SM measured -- and
        // holding all 48 plus the packed P made the dropout variant spill around its prefetched mask words):
        // pass 1 only takes the row maximum, pass 2 re-reads each chunk for the exponentials.
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent FMNMX3 chains
#pragma unroll
        for (int ch = 0; ch < HN / 16; ++ch) {
          uint32_t sr[16];
          tmem_ld_x16(s_addr + ch * 16, sr);
          tc_wait_ld();
          if (mb != 0ull) {
            const uint32_t m16 = static_cast<uint32_t>(mb >> (16 * ch)) & 0xffffu;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if ((m16 >> i) & 1u) sr[i] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int i = 0; i < 16; i += 8)
#pragma unroll
            for (int u = 0; u < 4; ++u)
              mx4[u] = max3(mx4[u], __uint_as_float(sr[i + 2 * u]), __uint_as_float(sr[i + 2 * u + 1]));
        }
        const float m_new = fmaxf(m_ref, fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * scale_log2);
        // lazy rescale: O and l stay relative to m_ref unless the max grew by more than 2^tau
        const bool grow = m_new > m_ref + lazy_tau;  // also true for the first finite tile (m_ref = -inf)
        const bool need = grow && (m_ref != -INFINITY);
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2_approx(m_ref - m_new) : 1.f;
          mbar_wait(&sm.o_full[hf], (f - 1) & 1, 8);  // P.V of the previous tile has landed
          tc_fence_after();
          uint32_t orr[32];
          tmem_ld_x32(tmem + lane_addr + C_O + hf * DH, orr);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < DH; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
          tmem_st_x32(tmem + lane_addr + C_O + hf * DH, orr);
          l *= alpha;
        }
        if (grow) m_ref = m_new;
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
        const uint64_t nm2 = pack_f32x2(-m_use, -m_use);
        const uint64_t db = hf == 0 ? (static_cast<uint64_t>(e0) | (static_cast<uint64_t>(e1 & 0xffffu) << 32))
                                    : (static_cast<uint64_t>(e0 >> 16) | (static_cast<uint64_t>(e1) << 16));
#pragma unroll
        for (int ch = 0; ch < HN / 16; ++ch) {
          uint32_t sr[16], pk[8];
          tmem_ld_x16(s_addr + ch * 16, sr);
          tc_wait_ld();
          if (mb != 0ull) {
            const uint32_t m16 = static_cast<uint32_t>(mb >> (16 * ch)) & 0xffffu;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if ((m16 >> i) & 1u) sr[i] = 0xff800000u;
          }
          const uint32_t d16 = DROP ? static_cast<uint32_t>(db >> (16 * ch)) & 0xffffu : 0u;
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // P column 8 ch + i holds keys 16 ch + 2i, 2i + 1 of the half
            const uint64_t x2 =
                fma_f32x2(pack_f32x2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), sc2, nm2);
            float x0, x1, p0, p1;
            unpack_f32x2(x2, x0, x1);
            if ((i & 3) < POLYQ) {
              ex2_poly_f32x2(x0, x1, p0, p1);
            } else {
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            rs2[i & 1] = add_f32x2(rs2[i & 1], pack_f32x2(p0, p1));  // the softmax denominator is NOT dropped
            if (DROP) {  // attention-probability dropout (nn.MultiheadAttention(dropout=p)): zero P, rescale O at the end
              if ((d16 >> (2 * i)) & 1u) p0 = 0.f;
              if ((d16 >> (2 * i + 1)) & 1u) p1 = 0.f;
            }
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st_x8(s_addr + ch * 8, pk);  // P over S columns this thread has already consumed
        }
        }
        float r0, r1, r2, r3;
        unpack_f32x2(rs2[0], r0, r1);
        unpack_f32x2(rs2[1], r2, r3);
        l += (r0 + r1) + (r2 + r3);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&sm.p_full[hf][f & 1]);
      }

      // ---- item epilogue: read O, free it for the next item, merge the two key halves ----
      // (o_full advances one phase per tile and is only waited on by the lazy rescale, where it is provably at
      // most one phase behind; the epilogue gets its own once-per-item barrier so parities cannot alias)
      mbar_wait(&sm.o_done[hf], it & 1, 9);
      tc_fence_after();
      uint32_t orr[32];
      tmem_ld_x32(tmem + lane_addr + C_O + hf * DH, orr);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&sm.o_free[hf]);
      // (32-bit shared addresses: a generic pointer into dynamic smem makes these generic LD/ST on the long path)
      const uint32_t xw = smem_u32(&sm.xch[hf][row][0]), xr = smem_u32(&sm.xch[1 - hf][row][0]);
      sts_u4(xw, __float_as_uint(m_ref), __float_as_uint(l), 0u, 0u);
#pragma unroll
      for (int i = 0; i < 16; i += 4)  // the 16 dims the other half finalises
        sts_u4(xw + 16 + i * 4, orr[(1 - hf) * 16 + i], orr[(1 - hf) * 16 + i + 1], orr[(1 - hf) * 16 + i + 2],
               orr[(1 - hf) * 16 + i + 3]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float4 ml = lds_f4(xr);
      const float m_o = ml.x, l_o = ml.y;
      const float m_all = fmaxf(m_ref, m_o);
      const float m_fin = (m_all == -INFINITY) ? 0.f : m_all;
      const float a_s = ex2_approx(m_ref - m_fin), a_o = ex2_approx(m_o - m_fin);
      const float l_all = l * a_s + l_o * a_o;
      const float inv = (DROP ? drop_scale(dp.thr16) : 1.f) / l_all;
      const float cs = a_s * inv, co = a_o * inv;
      const int qrow = qt * BM + row;
      uint32_t ob[8];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 t = lds_f4(xr + 16 + i * 4);
        ob[i / 2] = pack_bf16x2(__uint_as_float(orr[hf * 16 + i]) * cs + t.x * co,
                                __uint_as_float(orr[hf * 16 + i + 1]) * cs + t.y * co);
        ob[i / 2 + 1] = pack_bf16x2(__uint_as_float(orr[hf * 16 + i + 2]) * cs + t.z * co,
                                    __uint_as_float(orr[hf * 16 + i + 3]) * cs + t.w * co);
      }
      if (qrow < N) {
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b * N + qrow) * heads + h) * DH + hf * 16);
        dst[0] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
        dst[1] = make_uint4(ob[4], ob[5], ob[6], ob[7]);
        if (lse && hf == 0) lse[(static_cast<size_t>(b) * heads + h) * N + qrow] = m_all + lg2_approx(l_all);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // exchange buffer is reused by the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<256>(tmem);
}

}  // namespace
}  // namespace destr

extern "C" int destr_debug_knob(int idx, int value) {
  if (idx < 0 || idx >= 24) return 2;
  destr::g_knobs[idx] = value;
  return 0;
}

extern "C" int destr_enc_attn_fwd(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v,
                                  const uint32_t* mask_bits, int words_per_row, void* out, float* lse, int B, int N,
                                  int heads, float scale, const uint32_t* drop_rowbits, int drop_words,
                                  uint32_t drop_thr16, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(q && k && v && mask_bits && out, "null pointer");
  DESTR_CHECK_ARG(B > 0 && N > 0 && heads > 0 && heads * DH <= 256, "shape");
  DESTR_CHECK_ARG(words_per_row >= ceil_div(N, BN) * 3, "words_per_row (need >= 3 words per 96-key tile)");
  DESTR_CHECK_ARG(!drop_thr16 || (drop_rowbits && drop_words >= ceil_div(N, BN) * 3),
                  "dropout needs the row bit matrix of destr_attn_dropout_bits");
  const uint64_t rows = static_cast<uint64_t>(B) * N;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, q, rows, heads * DH, ld_q, BM, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, k, rows, heads * DH, ld_k, BN, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, v, rows, heads * DH, ld_v, BN, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  const size_t smem = sizeof(Smem) + 1024;
  using KernelT = decltype(&enc_attn_fwd_kernel<0, false>);
  static const KernelT kernels[8] = {enc_attn_fwd_kernel<0, false>, enc_attn_fwd_kernel<1, false>,
                                     enc_attn_fwd_kernel<2, false>, enc_attn_fwd_kernel<3, false>,
                                     enc_attn_fwd_kernel<0, true>,  enc_attn_fwd_kernel<1, true>,
                                     enc_attn_fwd_kernel<2, true>,  enc_attn_fwd_kernel<3, true>};
  for (int i = 0; i < 8; ++i) DESTR_SMEM_OPTIN(kernels[i], smem);
  const KernelT kernel = kernels[(g_knobs[9] & 3) + (drop_thr16 ? 4 : 0)];
  const int n_items = B * heads * ceil_div(N, BM);
  int grid = n_items < 2 * 148 ? n_items : 2 * 148;  // persistent: 2 CTAs per SM
  if (g_knobs[12] > 0 && g_knobs[12] < grid) grid = g_knobs[12];
  Knobs kn{(uint32_t)g_knobs[0], (uint32_t)g_knobs[1], (uint32_t)g_knobs[2],
           (uint32_t)g_knobs[3], (uint32_t)g_knobs[4], (uint32_t)g_knobs[5], (uint32_t)g_knobs[19]};
  kernel<<<grid, NTHREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      tq, tk, tv, mask_bits, words_per_row, static_cast<__nv_bfloat16*>(out), lse, N, heads, n_items,
      scale * 1.4426950408889634f, g_knobs[11] ? (float)(g_knobs[11] - 1) : kLazyTau, kn,
      DropBits{drop_rowbits, drop_thr16, drop_words});
  DESTR_LAUNCH_CHECK();
  return 0;
}
