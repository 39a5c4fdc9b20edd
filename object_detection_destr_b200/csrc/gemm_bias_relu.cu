// out = dropout(relu(A W^T + bias)), bf16 -- the first FFN layer of the encoder block and of the decoder's class / box
// branches (reference encoder_block.py:107-108 `dropout2(relu(fc1(x)))`, decoder_block.py:255), as ONE tcgen05 GEMM
// with the whole tail in its epilogue.  The library path was a cuBLASLt GEMM with a bias+ReLU epilogue followed by a
// separate in-place dropout pass over the [M, 2048] activation (68 MB of extra HBM traffic per encoder layer).
//   A   bf16 [M, K] row-major (K-major), row pitch lda;   W bf16 [N, K] row-major as nn.Linear stores it
//   out bf16 [M, N], row pitch ldo;   K == 256, N % 256 == 0; M arbitrary (TMA zero-fills, stores are guarded)
// With K = 256 the product has little arithmetic per operand byte: a 128x128 tile re-reads 128 KB of operands for
// 32 KB of output and the kernel is bound by L2 -> SM bandwidth (measured: 28 us).  So the WEIGHT block of a CTA
// (256 output features x 256 = 128 KB) stays resident in shared memory and only A streams: persistent CTAs, CTA =
// (n-block of 256 features, every gs-th 128-row tile), tile 128 x 256.  One warp = TMA producer (W once, then a ring
// of four 16 KB K-chunks of A, SWIZZLE_128B), one warp = MMA issuer (tcgen05.mma 128x256x16, fp32 accumulator in
// TMEM, two accumulators = all 512 columns, so the epilogue of tile i overlaps the mainloop of tile i+1), 16
// epilogue warps: thread = output row x 64-column group, 32 columns at a time out of TMEM, + bias (staged in smem per
// tile), ReLU, the counter-based dropout mask of common.cuh (the same function destr_dropout_inplace evaluates: one
// hash per column pair), bf16 pack, per-warp smem transpose, stores of 8 rows x 64 contiguous bytes per instruction.
// HBM-bound by the output write (34 MB at M = 8400, N = 2048): 2*M*N bytes out + 2*M*K in per n-block pass.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, KFIX = 256, NKC = KFIX / BK;
constexpr int NSTAGE = 4;
constexpr int NEPI = 16;                       // epilogue warps
constexpr int NTHREADS = (NEPI + 2) * 32;      // + TMA warp + MMA warp
constexpr uint32_t CHUNK_BYTES = BM * BK * 2;    // 16 KB: 128 rows x 128 B
constexpr uint32_t WCHUNK_BYTES = BN * BK * 2;   // 32 KB: 256 rows x 128 B


struct __align__(1024) Smem {
  uint8_t w[NKC][WCHUNK_BYTES];
  uint8_t a[NSTAGE][CHUNK_BYTES];
  float bias[BN];
  uint8_t stage[NEPI][32 * 64];  // epilogue: per-warp transpose tile
  uint64_t full[NSTAGE];
  uint64_t empty[NSTAGE];
  uint64_t w_full;
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_bias_relu_drop_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int M, int N,
                           int ldo, int relu, Drop dp) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (M + BM - 1) / BM, nb = N / BN;
  const int n0 = (static_cast<int>(blockIdx.x) % nb) * BN;           // this CTA's block of output features
  const int j0 = static_cast<int>(blockIdx.x) / nb, gs = static_cast<int>(gridDim.x) / nb;  // its m-tiles: j0, j0+gs, ..
  const int my_tiles = (mt - j0 + gs - 1) / gs;

  if (warp == NEPI && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.acc_full[s], 1);
      mbar_init(&sm.acc_empty[s], NEPI * 32);
    }
    mbar_init(&sm.w_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == NEPI + 1) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  pdl_wait();    // (common.cuh: programmatic dependent launch) prologue done, global memory is touched from here on
  pdl_launch();

  if (warp == NEPI) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      mbar_arrive_expect_tx(&sm.w_full, NKC * WCHUNK_BYTES);  // the resident weight block
      for (int kc = 0; kc < NKC; ++kc) tma_load_2d(sm.w[kc], &tm_w, &sm.w_full, kc * BK, n0);
      int c = 0;  // running chunk counter over all of this CTA's tiles
      for (int i = 0; i < my_tiles; ++i) {
        const int m0 = (j0 + i * gs) * BM;
        for (int kc = 0; kc < NKC; ++kc, ++c) {
          const int s = c % NSTAGE;
          mbar_wait(&sm.empty[s], ((c / NSTAGE) & 1) ^ 1, 41);
          mbar_arrive_expect_tx(&sm.full[s], CHUNK_BYTES);
          tma_load_2d(sm.a[s], &tm_a, &sm.full[s], kc * BK, m0);
        }
      }
    }
    __syncwarp();
  } else if (warp == NEPI + 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, false, false);
      constexpr uint64_t dconst = umma_desc_const(16, 1024, SWZ_128B);
      int c = 0;
      mbar_wait(&sm.w_full, 0, 45);
      for (int i = 0; i < my_tiles; ++i) {
        const int acc = i & 1;
        mbar_wait(&sm.acc_empty[acc], ((i >> 1) & 1) ^ 1, 42);  // the epilogue has drained this accumulator
        tc_fence_after();
        for (int kc = 0; kc < NKC; ++kc, ++c) {
          const int s = c % NSTAGE;
          mbar_wait(&sm.full[s], (c / NSTAGE) & 1, 43);
          tc_fence_after();
          const uint64_t da = dconst + (smem_u32(sm.a[s]) >> 4), db = dconst + (smem_u32(sm.w[kc]) >> 4);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks)  // 32 bytes (16 bf16) further along K inside the 128-byte swizzle row
            umma_ss(tmem + acc * BN, da + ks * 2, db + ks * 2, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
          tc_commit(&sm.empty[s]);
        }
        tc_commit(&sm.acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue warps ------------------------------
    // 16 warps: TMEM lane quadrant q = warp & 3 (a warp can only read lanes 32q..32q+31), column group cg = warp >> 2
    // (64 of the tile's 256 columns).  The per-element work is a dependent ALU chain, so it is the number of warps
    // per scheduler (4) that hides its latency: with 4 epilogue warps in all the kernel ran at IPC 0.25.
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int r = q * 32 + lane;  // row of the tile
    const uint32_t seed = (dp.thr16 && dp.seed) ? *dp.seed : 0u;
    const float ds = drop_scale(dp.thr16);
    // the CTA's feature block is fixed: its 256 bias values are staged once
    if (threadIdx.x < BN) sm.bias[threadIdx.x] = bias ? bias[n0 + threadIdx.x] : 0.f;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const uint32_t a_bias = smem_u32(&sm.bias[cg * 64]);
    const uint32_t stg = smem_u32(sm.stage[warp]);
    for (int i = 0; i < my_tiles; ++i) {
      const int m0 = (j0 + i * gs) * BM;
      const int acc = i & 1;
      mbar_wait(&sm.acc_full[acc], (i >> 1) & 1, 44);
      tc_fence_after();
      const int row = m0 + r;
      // both 32-column halves of this warp's group are fetched at once, and the accumulator is handed back to the MMA
      // warp before any arithmetic
      uint32_t v[2][32];
      tmem_ld_x32(tmem + lane_addr + acc * BN + cg * 64, v[0]);
      tmem_ld_x32(tmem + lane_addr + acc * BN + cg * 64 + 32, v[1]);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&sm.acc_empty[acc]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[16];
#pragma unroll
        for (int p4 = 0; p4 < 8; ++p4) {  // 4 columns per step: one 16-byte bias read
          const float4 bb = lds_f4(a_bias + (h * 32 + p4 * 4) * 4);
          const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int p = p4 * 2 + e;
            float x0 = __uint_as_float(v[h][2 * p]) + bv[2 * e];
            float x1 = __uint_as_float(v[h][2 * p + 1]) + bv[2 * e + 1];
            if (relu) {
              x0 = fmaxf(x0, 0.f);
              x1 = fmaxf(x1, 0.f);
            }
            if (dp.thr16) {
              const uint32_t bits = drop_bits(seed, dp.site, static_cast<uint32_t>(row), ((n0 + cg * 64 + h * 32) >> 1) + p);
              x0 = ((bits & 0xFFFFu) >= dp.thr16) ? x0 * ds : 0.f;
              x1 = ((bits >> 16) >= dp.thr16) ? x1 * ds : 0.f;
            }
            pk[p] = pack_bf16x2(x0, x1);
          }
        }
        // A thread owns a row, so a direct store instruction would touch 32 different lines with 16 bytes each (one
        // LSU transaction per lane: the 34 MB of output cost 2 M transactions and bound the kernel).  Transpose through
        // a per-warp tile (32 rows x 64 B, 16-byte chunks XOR-swizzled): each store instruction then writes 8 rows x
        // 64 contiguous bytes.
#pragma unroll
        for (int u = 0; u < 4; ++u)
          sts_u4(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rr = k * 8 + (lane >> 2), c = lane & 3;
          const float4 val = lds_f4(stg + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4));
          const int grow = m0 + q * 32 + rr;
          if (grow < M)
            *reinterpret_cast<float4*>(out + static_cast<size_t>(grow) * ldo + n0 + cg * 64 + h * 32 + c * 8) = val;
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NEPI + 1) tmem_dealloc<512>(tmem);
}

}  // namespace
}  // namespace destr

extern "C" int destr_linear_bias_relu_dropout(const void* a, int lda, const void* w, const float* bias, void* out,
                                              int ldo, int M, int N, int K, int relu, const uint32_t* drop_seed,
                                              uint32_t drop_thr16, uint32_t drop_site, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(a && w && out, "null pointer");
  DESTR_CHECK_ARG(M > 0 && N > 0 && K == KFIX && N % BN == 0, "shape (K == 256, N % 256 == 0)");
  DESTR_CHECK_ARG(lda % 8 == 0 && ldo % 8 == 0 && lda >= K && ldo >= N, "row pitches (multiples of 8 elements)");
  CUtensorMap ta, tw;
  int rc;
  if ((rc = make_tmap_bf16_2d(&ta, a, M, K, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tw, w, N, K, K, BN, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const size_t smem = sizeof(Smem) + 1024;
  DESTR_SMEM_OPTIN(gemm_bias_relu_drop_kernel, smem);
  const int mt = ceil_div(M, BM), nb = N / BN;
  int gs = 148 / nb;  // CTAs per n-block (each keeps that block's weights resident)
  if (gs < 1) gs = 1;
  if (gs > mt) gs = mt;
  // balance: no more CTAs per n-block than needed for ceil(mt / gs) tiles each
  gs = ceil_div(mt, ceil_div(mt, gs));
  const int grid = gs * nb;
  DESTR_CUDA(launch_k(gemm_bias_relu_drop_kernel, dim3(grid), dim3(NTHREADS), smem, static_cast<cudaStream_t>(stream), ta, tw,
                      bias, static_cast<__nv_bfloat16*>(out), M, N, ldo, relu, Drop{drop_seed, drop_thr16, drop_site}));
  return 0;
}
