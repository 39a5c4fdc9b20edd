// Encoder multi-head self-attention forward (d_head = 32), flash style, on tcgen05 / TMEM / TMA.
//
// Replaces the bmm -> softmax -> bmm chain of torch's F.multi_head_attention_forward as invoked by
// the reference at src/model/blocks/encoder_block.py:97-103 (need_weights=True math path: q*scale,
// QK^T, key-padding mask -> -inf, softmax, PV) without ever materialising the N x N scores.
//
// One CTA = 128 query rows of one (batch, head); 192 threads:
//   warps 0-3  softmax   (thread t <-> query row t <-> TMEM lane t; no cross-thread reductions)
//   warp  4    TMA producer  (Q once; K_j / V_j tiles through a KSTAGES ring)
//   warp  5    tcgen05.mma issuer + TMEM allocator
// Head split is done by TMA: Q/K/V tiles are 128 x 32 boxes of the token-major projection output
// (row pitch ld_*), 64-byte rows, SWIZZLE_64B.  Q,K are K-major UMMA operands; V is the MN-major B
// operand of P.V.  S_j = Q.K_j^T lands in one of TWO 128-column TMEM buffers, so Q.K_{j+1}^T is issued
// (and runs) while the softmax warps are still busy with S_j -- the exp (MUFU) pipe, not the tensor
// pipe, bounds this kernel at d_head = 32 (16 k exps vs 2 x 1 MFLOP per tile).  Softmax threads read S
// with tcgen05.ld, write P (bf16) back over the same columns with tcgen05.st, and the P.V MMA takes
// A = P straight from TMEM.  O_j = P_j.V_j lands in a double-buffered 32-column TMEM slot and is folded
// into a per-thread fp32 register accumulator with the online-softmax rescale.
//
// TMEM map (512 columns allocated): S0 [0,128) S1 [128,256) O0 [256,288) O1 [288,320).
// P_b aliases the first 64 columns of S_b.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {

int g_knobs[16] = {512, 512, 16, 512, 8, 1024, 16384, 1024, 2048, 0, 0, 0, 0, 0, 0, 0};

namespace {

constexpr int DH = 32;
constexpr int BM = 128;
constexpr int BN = 128;
constexpr int KSTAGES = 4;
constexpr int NTHREADS = 192;
constexpr uint32_t TILE_BYTES = BM * DH * 2;  // 8192

struct __align__(1024) Smem {
  uint8_t q[TILE_BYTES];
  uint8_t k[KSTAGES][TILE_BYTES];
  uint8_t v[KSTAGES][TILE_BYTES];
  uint64_t q_full;
  uint64_t kv_full[KSTAGES];
  uint64_t kv_empty[KSTAGES];
  uint64_t s_full[2];  // indexed by tile parity j & 1
  uint64_t p_full[2];
  uint64_t o_full[2];
  uint32_t tmem_base;
};

struct Knobs {
  uint32_t v_lbo, v_sbo, qk_lbo, qk_sbo, p_kstep_cols, v_kstep_bytes;
};

__global__ void __launch_bounds__(NTHREADS, 1)
enc_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const uint32_t* __restrict__ mask_bits,
                    int words_per_row, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int heads,
                    float scale_log2, Knobs kn) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BM;
  const int nkv = (N + BN - 1) / BN;
  const int row_base = b * N;  // first token row of this image

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(&sm.q_full, 1);
    for (int s = 0; s < KSTAGES; ++s) {
      mbar_init(&sm.kv_full[s], 1);
      mbar_init(&sm.kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.s_full[s], 1);
      mbar_init(&sm.p_full[s], BM);
      mbar_init(&sm.o_full[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 4) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      mbar_arrive_expect_tx(&sm.q_full, TILE_BYTES);
      tma_load_2d(sm.q, &tm_q, &sm.q_full, h * DH, row_base + q0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % KSTAGES;
        const uint32_t ph = (j / KSTAGES) & 1;
        mbar_wait(&sm.kv_empty[s], ph ^ 1, 1);
        mbar_arrive_expect_tx(&sm.kv_full[s], 2 * TILE_BYTES);
        tma_load_2d(sm.k[s], &tm_k, &sm.kv_full[s], h * DH, row_base + j * BN);
        tma_load_2d(sm.v[s], &tm_v, &sm.kv_full[s], h * DH, row_base + j * BN);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BM, BN, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BM, DH, false, true);
      auto issue_qk = [&](int j) {  // S_{j&1} = Q . K_j^T
        const int stage = j % KSTAGES;
        mbar_wait(&sm.kv_full[stage], (j / KSTAGES) & 1, 3);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          const uint64_t a = umma_smem_desc(smem_u32(sm.q) + ks * 32, kn.qk_lbo, kn.qk_sbo, SWZ_64B);
          const uint64_t bd = umma_smem_desc(smem_u32(sm.k[stage]) + ks * 32, kn.qk_lbo, kn.qk_sbo, SWZ_64B);
          umma_ss(tmem + (j & 1) * 128, a, bd, idesc_qk, ks > 0);
        }
        tc_commit(&sm.s_full[j & 1]);
      };
      mbar_wait(&sm.q_full, 0, 2);
      issue_qk(0);
      if (nkv > 1) issue_qk(1);
      for (int j = 0; j < nkv; ++j) {
        const int stage = j % KSTAGES;
        mbar_wait(&sm.p_full[j & 1], (j >> 1) & 1, 4);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {  // O_{j&1} = P_j . V_j
          const uint64_t bd =
              umma_smem_desc(smem_u32(sm.v[stage]) + ks * kn.v_kstep_bytes, kn.v_lbo, kn.v_sbo, SWZ_64B);
          umma_ts(tmem + 256 + (j & 1) * 32, tmem + (j & 1) * 128 + ks * kn.p_kstep_cols, bd, idesc_pv, ks > 0);
        }
        tc_commit(&sm.o_full[j & 1]);
        tc_commit(&sm.kv_empty[stage]);
        if (j + 2 < nkv) issue_qk(j + 2);  // reuses S_{j&1}: ordered behind P_j.V_j in the tensor pipe
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ softmax warps ------------------------------
    const int wq = warp;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t* mrow = mask_bits + static_cast<size_t>(b) * words_per_row;

    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    float acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const uint32_t s_addr = tmem + lane_addr + (j & 1) * 128;
      mbar_wait(&sm.s_full[j & 1], (j >> 1) & 1, 6);
      tc_fence_after();
      uint32_t sr[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(s_addr + c * 32, sr[c]);
      tc_wait_ld();

      const uint4 mw = *reinterpret_cast<const uint4*>(mrow + j * 4);
      const uint32_t mwa[4] = {mw.x, mw.y, mw.z, mw.w};
      if ((mw.x | mw.y | mw.z | mw.w) != 0u) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if ((mwa[c] >> i) & 1u) sr[c][i] = 0xff800000u;  // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains (ILP)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[c] = fmaxf(mx4[c], __uint_as_float(sr[c][i]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, mx * scale_log2);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2_approx(m - m_use);
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          // P columns [32c, 32c+32) hold keys [64c, 64c+64) packed two per column
          const int e = 2 * i;
          const float p0 = ex2_approx(fmaf(__uint_as_float(sr[2 * c + (e >> 5)][e & 31]), scale_log2, -m_use));
          const float p1 =
              ex2_approx(fmaf(__uint_as_float(sr[2 * c + ((e + 1) >> 5)][(e + 1) & 31]), scale_log2, -m_use));
          rs4[i & 3] += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_x32(s_addr + c * 32, pk);
      }
      l = l * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&sm.p_full[j & 1]);

      if (j > 0) {
        const int jb = (j - 1) & 1;
        mbar_wait(&sm.o_full[jb], ((j - 1) >> 1) & 1, 7);
        tc_fence_after();
        uint32_t orr[32];
        tmem_ld_x32(tmem + lane_addr + 256 + jb * 32, orr);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < DH; ++i) acc[i] = fmaf(acc[i], alpha_prev, __uint_as_float(orr[i]));
      }
      alpha_prev = alpha;
      m = m_new;
    }
    {
      const int jl = nkv - 1;
      mbar_wait(&sm.o_full[jl & 1], (jl >> 1) & 1, 8);
      tc_fence_after();
      uint32_t orr[32];
      tmem_ld_x32(tmem + lane_addr + 256 + (jl & 1) * 32, orr);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < DH; ++i) acc[i] = fmaf(acc[i], alpha_prev, __uint_as_float(orr[i]));
    }
    const int qrow = q0 + wq * 32 + lane;
    if (qrow < N) {
      const float inv = 1.f / l;
      uint32_t ob[DH / 2];
#pragma unroll
      for (int i = 0; i < DH / 2; ++i) ob[i] = pack_bf16x2(acc[2 * i] * inv, acc[2 * i + 1] * inv);
      uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(row_base + qrow) * heads + h) * DH);
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) dst[i] = make_uint4(ob[4 * i], ob[4 * i + 1], ob[4 * i + 2], ob[4 * i + 3]);
      if (lse) lse[(static_cast<size_t>(b) * heads + h) * N + qrow] = m + lg2_approx(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

}  // namespace
}  // namespace destr

extern "C" int destr_debug_knob(int idx, int value) {
  if (idx < 0 || idx >= 16) return 2;
  destr::g_knobs[idx] = value;
  return 0;
}

extern "C" int destr_enc_attn_fwd(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v,
                                  const uint32_t* mask_bits, int words_per_row, void* out, float* lse, int B, int N,
                                  int heads, float scale, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(q && k && v && mask_bits && out, "null pointer");
  DESTR_CHECK_ARG(B > 0 && N > 0 && heads > 0 && heads * DH <= 256, "shape");
  DESTR_CHECK_ARG(words_per_row >= ceil_div(N, BN) * 4 && (words_per_row % 4) == 0, "words_per_row");
  const uint64_t rows = static_cast<uint64_t>(B) * N;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, q, rows, heads * DH, ld_q, BM, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, k, rows, heads * DH, ld_k, BN, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, v, rows, heads * DH, ld_v, BN, DH, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  const size_t smem = sizeof(Smem) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    DESTR_CUDA(cudaFuncSetAttribute(enc_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  Knobs kn{(uint32_t)g_knobs[0], (uint32_t)g_knobs[1], (uint32_t)g_knobs[2],
           (uint32_t)g_knobs[3], (uint32_t)g_knobs[4], (uint32_t)g_knobs[5]};
  dim3 grid(ceil_div(N, BM), heads, B);
  enc_attn_fwd_kernel<<<grid, NTHREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      tq, tk, tv, mask_bits, words_per_row, static_cast<__nv_bfloat16*>(out), lse, N, heads,
      scale * 1.4426950408889634f, kn);
  DESTR_LAUNCH_CHECK();
  return 0;
}
