// tcgen05 GEMM family with fused epilogues: every projection / FFN / dX / dW product of the
// encoder (reference encoder_block.py:24-44, 88-112 and its autograd) runs through the kernels of this file instead
// of a library GEMM followed by separate elementwise / LayerNorm / column-sum passes.
//
//   C[M,N] = A[M,K] . op(B)      A bf16 row-major (K contiguous, pitch lda)
//                                 op(B): B_MN = false -> B is [N,K] row-major (nn.Linear weight: y = x W^T, forward)
//                                        B_MN = true  -> B is [K,N] row-major (the SAME weight used for dX = dY W:
//                                                        its N dimension is contiguous = "MN-major" UMMA operand)
// gemm_tc_kernel<BN, B_MN, EPI>: persistent CTAs, tile 128 x BN, K streamed in chunks of 64 through a TMA ring
// (A 16 KB + B BN*128 B per stage, SWIZZLE_128B), one TMA warp, one MMA warp (tcgen05.mma 128 x BN x 16, fp32
// accumulators double-buffered in TMEM so tile i's epilogue overlaps tile i+1's mainloop), 16 epilogue warps
// (thread = output row x BN/4 columns).  A CTA keeps ONE block of output columns (n0) for all its tiles, so the
// per-column vectors (bias, gamma, beta, column sums) are staged in shared memory once.
// Epilogues (all elementwise operands go through a per-warp shared-memory transpose so that every global access
// instruction touches 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes):
//   EPI_STORE     x = act(acc + bias) [dropout];  out = add + mul * x;  out2 = add2 + x      (each part optional)
//                 -> Linear(+ReLU)(+dropout), beta = 1 accumulation (dX + residual gradient), x + pos * s (the
//                    position-scale add, encoder_block.py:38,95) and its backward pair (ds = dxq * pos, dx += dxq)
//   EPI_RELU_BWD  dpre = scale * acc * (h > 0);  colsum[n] += sum_rows dpre      (backward of dropout(relu(fc1 x)):
//                 the dX GEMM of fc2, the ReLU/dropout mask and the bias gradient of fc1 in one kernel)
//   EPI_RES_LN    z = res + dropout(acc + bias);  y = LN(z);  [y2 = LN2(res2 + y)]   (N == BN == 256: a CTA owns whole
//                 rows; out-proj + dropout1 + norm1, and fc2 + dropout3 + norm2 + the encoder's shared norm,
//                 encoder_block.py:104-110, :40).  z (bf16) and the row statistics are what the LayerNorm backward reads.
// gemm_dw_kernel: dW[Nout,Kin] += dY[M,Nout]^T X[M,Kin]  (both operands MN-major), split over M across CTAs so a
// 256 x 256 weight gradient still fills the GPU (cuBLAS ran these on 32 CTAs), fp32 red.global.add into the flat fp32
// gradient buffer.
#include "../../include/destr_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace destr {
extern int g_knobs[24];
namespace {
#define g_dw_split (::destr::g_knobs[15])

constexpr int BM = 128, BK = 64;
constexpr int NEPI = 16;
constexpr int NTHREADS = (NEPI + 2) * 32;
constexpr uint32_t A_BYTES = BM * BK * 2;  // 16 KB
constexpr float kEps = 1e-5f;

enum { EPI_STORE = 0, EPI_RELU_BWD = 1, EPI_RES_LN = 2 };

struct GemmArgs {
  int M, N, K;
  // batched products (blockIdx.y = batch index): A and B are stacks of per-batch row blocks inside ONE 2-D tensor map
  // (a_brows / b_brows rows per batch); `out` is the stack of the [M, N] results.  0 = not batched.
  int a_brows, b_brows;
  const float* bias;
  int relu;
  Drop dp;
  const __nv_bfloat16* mul;
  int ldmul;
  const __nv_bfloat16* add;
  int ldadd;
  __nv_bfloat16* out;
  int ldo;
  const __nv_bfloat16* add2;
  int ldadd2;
  __nv_bfloat16* out2;
  int ldo2;
  // EPI_RELU_BWD
  const __nv_bfloat16* hmask;
  int ldh;
  float scale;
  float* colsum;
  // EPI_RES_LN
  const __nv_bfloat16* res;
  int ldres;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* z;
  int ldz;
  float* mean;
  float* rstd;
  const __nv_bfloat16* res2;
  int ldres2;
  const float* gamma2;
  const float* beta2;
  __nv_bfloat16* y2;
  int ldy2;
  float* mean2;
  float* rstd2;
};

template <int BN>
struct __align__(1024) Smem {
  static constexpr int NSTAGE = BN == 256 ? 3 : 4;
  uint8_t a[NSTAGE][A_BYTES];
  uint8_t b[NSTAGE][BN * BK * 2];
  uint8_t stage[NEPI][32 * 64];  // per-warp transpose tile: 32 rows x 64 B
  float vec[6][BN];              // bias | gamma | beta | colsum | gamma2 | beta2
  float part[4][4][BM];          // LayerNorm partial row statistics [phase][column group][row]
  uint64_t full[NSTAGE];
  uint64_t empty[NSTAGE];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// Warp-cooperative load of a 32-row x 32-column bf16 tile (64 B per row) at (row0, col0): coalesced 8 rows x 64 B per
// instruction into the warp's staging tile, then every lane reads ITS row (row0 + lane) back: o[j] = columns 2j, 2j+1.
// Rows >= M read as zero.
__device__ __forceinline__ void warp_load_tile(uint32_t stg, const __nv_bfloat16* g, int ld, int row0, int col0, int M,
                                               int lane, uint32_t (&o)[16]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = k * 8 + (lane >> 2), c = lane & 3;
    const int grow = row0 + rr;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (grow < M) val = *reinterpret_cast<const uint4*>(g + static_cast<size_t>(grow) * ld + col0 + c * 8);
    sts_u4(stg + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4), val.x, val.y, val.z, val.w);
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float4 f = lds_f4(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4));
    o[4 * u] = __float_as_uint(f.x);
    o[4 * u + 1] = __float_as_uint(f.y);
    o[4 * u + 2] = __float_as_uint(f.z);
    o[4 * u + 3] = __float_as_uint(f.w);
  }
  __syncwarp();
}
// The mirror: every lane hands in its row (16 packed bf16 pairs), the warp stores 8 rows x 64 B per instruction.
__device__ __forceinline__ void warp_store_tile(uint32_t stg, __nv_bfloat16* g, int ld, int row0, int col0, int M,
                                                int lane, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int u = 0; u < 4; ++u)
    sts_u4(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = k * 8 + (lane >> 2), c = lane & 3;
    const float4 val = lds_f4(stg + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4));
    const int grow = row0 + rr;
    if (grow < M) *reinterpret_cast<float4*>(g + static_cast<size_t>(grow) * ld + col0 + c * 8) = val;
  }
  __syncwarp();
}

// sum over the warp's 32 rows of 32 per-lane column values: afterwards lane l holds the total of column l in v[0]
// (butterfly transpose-reduction: 31 shuffles instead of 32 x 5)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// one K chunk of a tile into ring slot c % NSTAGE (the tensor maps are passed as pointers to the kernel's
// __grid_constant__ parameters: TMA must read the descriptor from param/const/global space, never from a local copy)
template <int BN, bool B_MN>
__device__ __forceinline__ void issue_chunk(Smem<BN>& sm, const CUtensorMap* ta, const CUtensorMap* tb, int c, int m0,
                                            int kc, int n0, int a_row0, int b_row0) {
  constexpr int NSTAGE = Smem<BN>::NSTAGE;
  const int s = c % NSTAGE;
  mbar_arrive_expect_tx(&sm.full[s], A_BYTES + BN * BK * 2);
  tma_load_2d(sm.a[s], ta, &sm.full[s], kc * BK, a_row0 + m0);
  if (!B_MN) {
    tma_load_2d(sm.b[s], tb, &sm.full[s], kc * BK, b_row0 + n0);  // [BN rows (n)] x [64 k]: K-major
  } else {
#pragma unroll
    for (int a64 = 0; a64 < BN / 64; ++a64)  // [64 k rows] x [64 n]: one MN-major SW128 atom column per box
      tma_load_2d(sm.b[s] + a64 * 8192, tb, &sm.full[s], n0 + a64 * 64, b_row0 + kc * BK);
  }
}

template <int BN, bool B_MN, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_b1,
               const __grid_constant__ GemmArgs ga1, const __grid_constant__ CUtensorMap tm_a2,
               const __grid_constant__ CUtensorMap tm_b2, const __grid_constant__ GemmArgs ga2) {
  // blockIdx.z selects one of two independent problems of the same (N, K) family sharing the launch (e.g. dq_obj and
  // dq_pos of the split cross-attention backward); ordinary launches have gridDim.z = 1 and the second set unused
  // (the tensor maps are only ever addressed directly as kernel parameters -- TMA must read a descriptor from param /
  // const / global space, and a run-time selected pointer makes the compiler copy them to the local stack; the plain
  // argument struct is simply copied)
  const bool second = blockIdx.z != 0;
  const GemmArgs ga = second ? ga2 : ga1;
  using SM = Smem<BN>;
  constexpr int NSTAGE = SM::NSTAGE;
  constexpr int CW = BN / 4;   // output columns per epilogue warp
  constexpr int NH = CW / 32;  // 32-column halves per warp
  extern __shared__ uint8_t smem_raw[];
  SM& sm = *reinterpret_cast<SM*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = ga.M, N = ga.N;
  const int mt = (M + BM - 1) / BM, nb = (N + BN - 1) / BN;
  const int nkc = (ga.K + BK - 1) / BK;
  const int n0 = (static_cast<int>(blockIdx.x) % nb) * BN;
  const int j0 = static_cast<int>(blockIdx.x) / nb, gs = static_cast<int>(gridDim.x) / nb;
  const int my_tiles = (mt - j0 + gs - 1) / gs;
  const int bz = blockIdx.y;  // batch index (0 when not batched)

  if (warp == NEPI && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.acc_full[s], 1);
      mbar_init(&sm.acc_empty[s], NEPI * 32);
    }
    fence_mbar_init();
    if (second) {
      tma_prefetch_desc(&tm_a2);
      tma_prefetch_desc(&tm_b2);
    } else {
      tma_prefetch_desc(&tm_a1);
      tma_prefetch_desc(&tm_b1);
    }
  }
  if (warp == NEPI + 1) tmem_alloc<2 * BN>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  pdl_wait();    // prologue done; from here on global memory is touched: the previous kernel must have completed
  pdl_launch();  // and the next kernel may start its own prologue

  if (warp == NEPI) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int c = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m0 = (j0 + i * gs) * BM;
        for (int kc = 0; kc < nkc; ++kc, ++c) {
          mbar_wait(&sm.empty[c % NSTAGE], ((c / NSTAGE) & 1) ^ 1, 51);
          if (second)
            issue_chunk<BN, B_MN>(sm, &tm_a2, &tm_b2, c, m0, kc, n0, bz * ga.a_brows, bz * ga.b_brows);
          else
            issue_chunk<BN, B_MN>(sm, &tm_a1, &tm_b1, c, m0, kc, n0, bz * ga.a_brows, bz * ga.b_brows);
        }
      }
    }
    __syncwarp();
  } else if (warp == NEPI + 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, false, B_MN);
      constexpr uint64_t D_K = umma_desc_const(16, 1024, SWZ_128B);      // K-major: 8-row groups 1024 B apart
      constexpr uint64_t D_MN = umma_desc_const(8192, 1024, SWZ_128B);   // MN-major: 64-column atoms 8192 B apart
      int c = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int acc = i & 1;
        mbar_wait(&sm.acc_empty[acc], ((i >> 1) & 1) ^ 1, 52);
        tc_fence_after();
        for (int kc = 0; kc < nkc; ++kc, ++c) {
          const int s = c % NSTAGE;
          mbar_wait(&sm.full[s], (c / NSTAGE) & 1, 53);
          tc_fence_after();
          const uint64_t da = D_K + (smem_u32(sm.a[s]) >> 4);
          const uint64_t db = (B_MN ? D_MN : D_K) + (smem_u32(sm.b[s]) >> 4);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks)  // K-major: 32 B along the swizzled row; MN-major: 16 k-rows = 2048 B
            umma_ss(tmem + acc * BN, da + ks * 2, db + (B_MN ? ks * 128 : ks * 2), idesc, (kc > 0 || ks > 0) ? 1u : 0u);
          tc_commit(&sm.empty[s]);
        }
        tc_commit(&sm.acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue warps ------------------------------
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int r = q * 32 + lane;
    const uint32_t seed = (ga.dp.thr16 && ga.dp.seed) ? *ga.dp.seed : 0u;
    const float ds = drop_scale(ga.dp.thr16);
    const uint32_t thr16 = ga.dp.thr16;
    if (threadIdx.x < BN) {
      const int n = n0 + threadIdx.x;
      const bool in = n < N;
      sm.vec[0][threadIdx.x] = (ga.bias && in) ? ga.bias[n] : 0.f;
      sm.vec[3][threadIdx.x] = 0.f;
      if (EPI == EPI_RES_LN) {
        sm.vec[1][threadIdx.x] = in ? ga.gamma[n] : 0.f;
        sm.vec[2][threadIdx.x] = in ? ga.beta[n] : 0.f;
        sm.vec[4][threadIdx.x] = (ga.res2 && in) ? ga.gamma2[n] : 0.f;
        sm.vec[5][threadIdx.x] = (ga.res2 && in) ? ga.beta2[n] : 0.f;
      }
    }
    epi_bar();
    const uint32_t stg = smem_u32(sm.stage[warp]);
    const int cw0 = cg * CW;  // first column of this warp inside the tile
    __nv_bfloat16* const out_b = ga.out + static_cast<size_t>(bz) * M * ga.ldo;  // batched: stack of [M, N] results
    for (int i = 0; i < my_tiles; ++i) {
      const int m0 = (j0 + i * gs) * BM;
      const int acc = i & 1;
      const int row0 = m0 + q * 32;  // first row of this warp
      const int row = m0 + r;
      if (EPI != EPI_RES_LN) {
        mbar_wait(&sm.acc_full[acc], (i >> 1) & 1, 54);
        tc_fence_after();
      }

      if (EPI == EPI_RES_LN) {
        // ---- z = res + dropout(acc + bias); y = LN(z); optionally y2 = LN2(res2 + y) ----
        // z replaces the accumulator IN TMEM (fp32), so only 32 columns are in registers at any time (16 epilogue
        // warps leave 96 registers per thread); the statistics are two-pass (mean, then centred squares) like
        // destr_add_layernorm_fwd.  The accumulator returns to the MMA warp when the tile is finished.
        const uint32_t tz = tmem + lane_addr + acc * BN + cw0;
        float s1 = 0.f;
        // the residual tile is fetched while the mainloop of this tile is still running
        uint32_t rsa[NH][16];
#pragma unroll
        for (int h = 0; h < NH; ++h) warp_load_tile(stg, ga.res, ga.ldres, row0, n0 + cw0 + h * 32, M, lane, rsa[h]);
        mbar_wait(&sm.acc_full[acc], (i >> 1) & 1, 54);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          uint32_t v[32];
          uint32_t(&rs)[16] = rsa[h];
          tmem_ld_x32(tz + h * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int p = 0; p < 16; ++p) {
            float x0 = __uint_as_float(v[2 * p]) + sm.vec[0][cw0 + h * 32 + 2 * p];
            float x1 = __uint_as_float(v[2 * p + 1]) + sm.vec[0][cw0 + h * 32 + 2 * p + 1];
            if (thr16) {
              const uint32_t bits = drop_bits(seed, ga.dp.site, static_cast<uint32_t>(row), ((n0 + cw0 + h * 32) >> 1) + p);
              x0 = ((bits & 0xFFFFu) >= thr16) ? x0 * ds : 0.f;
              x1 = ((bits >> 16) >= thr16) ? x1 * ds : 0.f;
            }
            x0 += bf_lo(rs[p]);
            x1 += bf_hi(rs[p]);
            v[2 * p] = __float_as_uint(x0);
            v[2 * p + 1] = __float_as_uint(x1);
            s1 += x0 + x1;
          }
          tmem_st_x32(tz + h * 32, v);
          if (ga.z) {  // pre-LayerNorm sum (bf16): what the LayerNorm backward recomputes xhat from
#pragma unroll
            for (int p = 0; p < 16; ++p) rs[p] = pack_bf16x2(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1]));
            warp_store_tile(stg, ga.z, ga.ldz, row0, n0 + cw0 + h * 32, M, lane, rs);
          }
        }
        tc_wait_st();
        sm.part[0][cg][r] = s1;
        epi_bar();
        const float mean = (sm.part[0][0][r] + sm.part[0][1][r] + sm.part[0][2][r] + sm.part[0][3][r]) * (1.f / BN);
        float s2 = 0.f;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          uint32_t v[32];
          tmem_ld_x32(tz + h * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(v[j]) - mean;
            s2 = fmaf(d, d, s2);
          }
        }
        sm.part[1][cg][r] = s2;
        epi_bar();
        const float rstd =
            rsqrtf((sm.part[1][0][r] + sm.part[1][1][r] + sm.part[1][2][r] + sm.part[1][3][r]) * (1.f / BN) + kEps);
        if (cg == 0 && row < M) {
          ga.mean[row] = mean;
          ga.rstd[row] = rstd;
        }
        float s3 = 0.f;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          uint32_t v[32], pk[16];
          tmem_ld_x32(tz + h * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int p = 0; p < 16; ++p) {
            const int cc = cw0 + h * 32 + 2 * p;
            const float y0 = fmaf((__uint_as_float(v[2 * p]) - mean) * rstd, sm.vec[1][cc], sm.vec[2][cc]);
            const float y1 = fmaf((__uint_as_float(v[2 * p + 1]) - mean) * rstd, sm.vec[1][cc + 1], sm.vec[2][cc + 1]);
            pk[p] = pack_bf16x2(y0, y1);
          }
          warp_store_tile(stg, ga.out, ga.ldo, row0, n0 + cw0 + h * 32, M, lane, pk);
          if (ga.res2) {  // second LayerNorm on res2 + y (y as stored: bf16); its input goes back to TMEM
            uint32_t rs[16];
            warp_load_tile(stg, ga.res2, ga.ldres2, row0, n0 + cw0 + h * 32, M, lane, rs);
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const float a0 = bf_lo(pk[p]) + bf_lo(rs[p]), a1 = bf_hi(pk[p]) + bf_hi(rs[p]);
              v[2 * p] = __float_as_uint(a0);
              v[2 * p + 1] = __float_as_uint(a1);
              s3 += a0 + a1;
            }
            tmem_st_x32(tz + h * 32, v);
          }
        }
        if (ga.res2) {  // uniform over the CTA
          tc_wait_st();
          sm.part[2][cg][r] = s3;
          epi_bar();
          const float mean2 = (sm.part[2][0][r] + sm.part[2][1][r] + sm.part[2][2][r] + sm.part[2][3][r]) * (1.f / BN);
          float s4 = 0.f;
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            uint32_t v[32];
            tmem_ld_x32(tz + h * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = __uint_as_float(v[j]) - mean2;
              s4 = fmaf(d, d, s4);
            }
          }
          sm.part[3][cg][r] = s4;
          epi_bar();
          const float rstd2 =
              rsqrtf((sm.part[3][0][r] + sm.part[3][1][r] + sm.part[3][2][r] + sm.part[3][3][r]) * (1.f / BN) + kEps);
          if (cg == 0 && row < M) {
            ga.mean2[row] = mean2;
            ga.rstd2[row] = rstd2;
          }
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            uint32_t v[32], pk[16];
            tmem_ld_x32(tz + h * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const int cc = cw0 + h * 32 + 2 * p;
              const float y0 = fmaf((__uint_as_float(v[2 * p]) - mean2) * rstd2, sm.vec[4][cc], sm.vec[5][cc]);
              const float y1 = fmaf((__uint_as_float(v[2 * p + 1]) - mean2) * rstd2, sm.vec[4][cc + 1], sm.vec[5][cc + 1]);
              pk[p] = pack_bf16x2(y0, y1);
            }
            warp_store_tile(stg, ga.y2, ga.ldy2, row0, n0 + cw0 + h * 32, M, lane, pk);
          }
        }
        tc_fence_before();
        mbar_arrive(&sm.acc_empty[acc]);
      } else {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          uint32_t v[32];
          tmem_ld_x32(tmem + lane_addr + acc * BN + cw0 + h * 32, v);
          tc_wait_ld();
          if (h == NH - 1) {  // the accumulator goes back to the MMA warp as soon as it is in registers
            tc_fence_before();
            mbar_arrive(&sm.acc_empty[acc]);
          }
          const int col0 = n0 + cw0 + h * 32;
          if (col0 >= N) continue;
          uint32_t pk[16], t16[16];
          if (EPI == EPI_RELU_BWD) {
            warp_load_tile(stg, ga.hmask, ga.ldh, row0, col0, M, lane, t16);
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const float x0 = bf_lo(t16[p]) > 0.f ? __uint_as_float(v[2 * p]) * ga.scale : 0.f;
              const float x1 = bf_hi(t16[p]) > 0.f ? __uint_as_float(v[2 * p + 1]) * ga.scale : 0.f;
              pk[p] = pack_bf16x2(x0, x1);
            }
            warp_store_tile(stg, ga.out, ga.ldo, row0, col0, M, lane, pk);
            if (ga.colsum) {  // bias gradient: column sums of what was stored (rows >= M are zero: their mask is)
              float x[32];
#pragma unroll
              for (int p = 0; p < 16; ++p) {
                x[2 * p] = bf_lo(pk[p]);
                x[2 * p + 1] = bf_hi(pk[p]);
              }
              const float tot = warp_colsum32(x, lane);
              atomicAdd(&sm.vec[3][cw0 + h * 32 + lane], tot);
            }
          } else {
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              float x0 = __uint_as_float(v[2 * p]) + sm.vec[0][cw0 + h * 32 + 2 * p];
              float x1 = __uint_as_float(v[2 * p + 1]) + sm.vec[0][cw0 + h * 32 + 2 * p + 1];
              if (ga.relu) {
                x0 = fmaxf(x0, 0.f);
                x1 = fmaxf(x1, 0.f);
              }
              if (thr16) {
                const uint32_t bits = drop_bits(seed, ga.dp.site, static_cast<uint32_t>(row), (col0 >> 1) + p);
                x0 = ((bits & 0xFFFFu) >= thr16) ? x0 * ds : 0.f;
                x1 = ((bits >> 16) >= thr16) ? x1 * ds : 0.f;
              }
              v[2 * p] = __float_as_uint(x0);
              v[2 * p + 1] = __float_as_uint(x1);
            }
            if (ga.out2) {  // out2 = add2 + x (before mul / add)
              if (ga.add2) warp_load_tile(stg, ga.add2, ga.ldadd2, row0, col0, M, lane, t16);
#pragma unroll
              for (int p = 0; p < 16; ++p) {
                float x0 = __uint_as_float(v[2 * p]), x1 = __uint_as_float(v[2 * p + 1]);
                if (ga.add2) {
                  x0 += bf_lo(t16[p]);
                  x1 += bf_hi(t16[p]);
                }
                pk[p] = pack_bf16x2(x0, x1);
              }
              warp_store_tile(stg, ga.out2, ga.ldo2, row0, col0, M, lane, pk);
            }
            if (ga.mul) {
              warp_load_tile(stg, ga.mul, ga.ldmul, row0, col0, M, lane, t16);
#pragma unroll
              for (int p = 0; p < 16; ++p) {
                v[2 * p] = __float_as_uint(__uint_as_float(v[2 * p]) * bf_lo(t16[p]));
                v[2 * p + 1] = __float_as_uint(__uint_as_float(v[2 * p + 1]) * bf_hi(t16[p]));
              }
            }
            if (ga.add) {
              warp_load_tile(stg, ga.add, ga.ldadd, row0, col0, M, lane, t16);
#pragma unroll
              for (int p = 0; p < 16; ++p) {
                v[2 * p] = __float_as_uint(__uint_as_float(v[2 * p]) + bf_lo(t16[p]));
                v[2 * p + 1] = __float_as_uint(__uint_as_float(v[2 * p + 1]) + bf_hi(t16[p]));
              }
            }
#pragma unroll
            for (int p = 0; p < 16; ++p) pk[p] = pack_bf16x2(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1]));
            warp_store_tile(stg, out_b, ga.ldo, row0, col0, M, lane, pk);
          }
        }
      }
    }
    if (EPI == EPI_RELU_BWD && ga.colsum) {
      epi_bar();
      if (threadIdx.x < BN && n0 + static_cast<int>(threadIdx.x) < N)
        atomicAdd(ga.colsum + n0 + threadIdx.x, sm.vec[3][threadIdx.x]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NEPI + 1) tmem_dealloc<2 * BN>(tmem);
}

// ------------------------------------------------------------------------------------------------------
// dW[Nout,Kin] += dY[M,Nout]^T X[M,Kin]: CTA = (128 x 128 output tile, split of the M rows); both operands are read
// as MN-major SW128 tiles straight from their row-major layout (no transposes), fp32 red.global.add at the end.
// ------------------------------------------------------------------------------------------------------
constexpr int DW_BN = 128, DW_STAGES = 3, DW_THREADS = 192;
struct __align__(1024) SmemDw {
  uint8_t a[DW_STAGES][BM * BK * 2];
  uint8_t b[DW_STAGES][DW_BN * BK * 2];
  uint64_t full[DW_STAGES];
  uint64_t empty[DW_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(DW_THREADS, 2)
gemm_dw_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
               float* __restrict__ dw, int lddw, int Nout, int Kin, int nchunk, int per_split) {
  extern __shared__ uint8_t smem_raw[];
  SmemDw& sm = *reinterpret_cast<SmemDw*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ktiles = (Kin + DW_BN - 1) / DW_BN;
  const int n0 = (static_cast<int>(blockIdx.x) / ktiles) * BM, k0 = (static_cast<int>(blockIdx.x) % ktiles) * DW_BN;
  const int c_begin = blockIdx.y * per_split;
  const int c_end = min(nchunk, c_begin + per_split);
  const int nc = c_end - c_begin;  // >= 1 by construction of the grid

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < DW_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 5) tmem_alloc<DW_BN>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  pdl_wait();
  pdl_launch();

  if (warp == 4) {
    if (elect_one()) {
      for (int c = 0; c < nc; ++c) {
        const int s = c % DW_STAGES;
        const int m = (c_begin + c) * BK;
        mbar_wait(&sm.empty[s], ((c / DW_STAGES) & 1) ^ 1, 61);
        mbar_arrive_expect_tx(&sm.full[s], BM * BK * 2 + DW_BN * BK * 2);
#pragma unroll
        for (int a64 = 0; a64 < BM / 64; ++a64) tma_load_2d(sm.a[s] + a64 * 8192, &tm_dy, &sm.full[s], n0 + a64 * 64, m);
#pragma unroll
        for (int a64 = 0; a64 < DW_BN / 64; ++a64) tma_load_2d(sm.b[s] + a64 * 8192, &tm_x, &sm.full[s], k0 + a64 * 64, m);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, DW_BN, true, true);
      constexpr uint64_t D_MN = umma_desc_const(8192, 1024, SWZ_128B);
      for (int c = 0; c < nc; ++c) {
        const int s = c % DW_STAGES;
        mbar_wait(&sm.full[s], (c / DW_STAGES) & 1, 62);
        tc_fence_after();
        const uint64_t da = D_MN + (smem_u32(sm.a[s]) >> 4), db = D_MN + (smem_u32(sm.b[s]) >> 4);
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) umma_ss(tmem, da + ks * 128, db + ks * 128, idesc, (c > 0 || ks > 0) ? 1u : 0u);
        tc_commit(&sm.empty[s]);
      }
      tc_commit(&sm.acc_full);
    }
    __syncwarp();
  } else {
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const int n = n0 + warp * 32 + lane;
    mbar_wait(&sm.acc_full, 0, 63);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < DW_BN / 32; ++h) {
      uint32_t v[32];
      tmem_ld_x32(tmem + lane_addr + h * 32, v);
      tc_wait_ld();
      if (n < Nout) {
        float* dst = dw + static_cast<size_t>(n) * lddw + k0 + h * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (k0 + h * 32 + j * 4 < Kin)
            red_add_v4(dst + j * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<DW_BN>(tmem);
}

template <int BN, bool B_MN>
int make_maps(const void* a, int lda, const void* b, int ldb, const GemmArgs& ga, int batch, CUtensorMap* ta,
              CUtensorMap* tb) {
  int rc;
  const uint64_t a_rows = batch > 1 ? static_cast<uint64_t>(batch) * ga.a_brows : ga.M;
  const uint64_t b_rows = batch > 1 ? static_cast<uint64_t>(batch) * ga.b_brows : (B_MN ? ga.K : ga.N);
  if ((rc = make_tmap_bf16_2d(ta, a, a_rows, ga.K, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if (!B_MN) {
    if ((rc = make_tmap_bf16_2d(tb, b, b_rows, ga.K, ldb, BN, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  } else {
    if ((rc = make_tmap_bf16_2d(tb, b, b_rows, ga.N, ldb, BK, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  return 0;
}

// `second` (optional): another problem with the same N, K, batch in the same launch (gridDim.z = 2)
template <int BN, bool B_MN, int EPI>
int launch_gemm(const void* a, int lda, const void* b, int ldb, const GemmArgs& ga, cudaStream_t st, int batch = 1,
                const void* a2 = nullptr, int lda2 = 0, const void* b2 = nullptr, int ldb2 = 0,
                const GemmArgs* second = nullptr) {
  CUtensorMap ta, tb, ta2, tb2;
  int rc;
  if ((rc = make_maps<BN, B_MN>(a, lda, b, ldb, ga, batch, &ta, &tb))) return rc;
  GemmArgs g2 = ga;
  if (second) {
    g2 = *second;
    if ((rc = make_maps<BN, B_MN>(a2, lda2, b2, ldb2, g2, batch, &ta2, &tb2))) return rc;
  } else {
    ta2 = ta;
    tb2 = tb;
  }
  const size_t smem = sizeof(Smem<BN>) + 1024;
  DESTR_SMEM_OPTIN((gemm_tc_kernel<BN, B_MN, EPI>), smem);
  const int mmax = second && second->M > ga.M ? second->M : ga.M;
  const int mt = ceil_div(mmax, BM), nb = ceil_div(ga.N, BN);
  int gs = 148 / nb;
  if (gs < 1) gs = 1;
  if (gs > mt) gs = mt;
  gs = ceil_div(mt, ceil_div(mt, gs));
  DESTR_CUDA(launch_k(gemm_tc_kernel<BN, B_MN, EPI>, dim3(gs * nb, batch, second ? 2 : 1), dim3(NTHREADS), smem, st, ta, tb,
                      ga, ta2, tb2, g2));
  return 0;
}

int check_common(const void* a, int lda, const void* b, int ldb, int M, int N, int K, int b_kn) {
  DESTR_CHECK_ARG(a && b, "null pointer");
  DESTR_CHECK_ARG(M > 0 && N > 0 && K > 0 && N % 32 == 0 && K % 8 == 0, "shape (N % 32 == 0, K % 8 == 0)");
  DESTR_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= (b_kn ? N : K), "row pitches (multiples of 8 elements)");
  return 0;
}

}  // namespace
}  // namespace destr

using namespace destr;
typedef __nv_bfloat16 bf16;

extern "C" int destr_gemm_bf16(const void* a, int lda, const void* b, int ldb, int b_kn, int M, int N, int K,
                               const float* bias, int relu, const uint32_t* drop_seed, uint32_t drop_thr16,
                               uint32_t drop_site, const void* mul, int ldmul, const void* add, int ldadd, void* out,
                               int ldo, const void* add2, int ldadd2, void* out2, int ldo2, void* stream) {
  int rc = check_common(a, lda, b, ldb, M, N, K, b_kn);
  if (rc) return rc;
  DESTR_CHECK_ARG(out && ldo % 8 == 0 && ldo >= N, "out / ldo");
  DESTR_CHECK_ARG((!mul || ldmul % 8 == 0) && (!add || ldadd % 8 == 0) && (!add2 || ldadd2 % 8 == 0) &&
                      (!out2 || (ldo2 % 8 == 0 && ldo2 >= N)), "epilogue operand pitches");
  GemmArgs ga{};
  ga.M = M, ga.N = N, ga.K = K;
  ga.bias = bias, ga.relu = relu, ga.dp = Drop{drop_seed, drop_thr16, drop_site};
  ga.mul = static_cast<const bf16*>(mul), ga.ldmul = ldmul;
  ga.add = static_cast<const bf16*>(add), ga.ldadd = ldadd;
  ga.out = static_cast<bf16*>(out), ga.ldo = ldo;
  ga.add2 = static_cast<const bf16*>(add2), ga.ldadd2 = ldadd2;
  ga.out2 = static_cast<bf16*>(out2), ga.ldo2 = ldo2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 128-wide tiles when the 256-wide ones would leave most SMs without a tile (knob 17: 1 / 2 force 128 / 256)
  bool narrow = ceil_div(M, BM) * ceil_div(N, 256) < 100 || N % 256 != 0;
  if (g_knobs[17] == 1) narrow = true;
  if (g_knobs[17] == 2 && N % 256 == 0) narrow = false;
  if (b_kn) return narrow ? launch_gemm<128, true, EPI_STORE>(a, lda, b, ldb, ga, st)
                          : launch_gemm<256, true, EPI_STORE>(a, lda, b, ldb, ga, st);
  return narrow ? launch_gemm<128, false, EPI_STORE>(a, lda, b, ldb, ga, st)
                : launch_gemm<256, false, EPI_STORE>(a, lda, b, ldb, ga, st);
}

extern "C" int destr_gemm_relu_bwd(const void* dy, int lddy, const void* w, int ldw, int M, int N, int K, const void* h,
                                   int ldh, float scale, void* dpre, int ldo, float* dbias, void* stream) {
  int rc = check_common(dy, lddy, w, ldw, M, N, K, 1);
  if (rc) return rc;
  DESTR_CHECK_ARG(h && dpre && ldh % 8 == 0 && ldo % 8 == 0 && ldh >= N && ldo >= N, "h / dpre");
  GemmArgs ga{};
  ga.M = M, ga.N = N, ga.K = K;
  ga.hmask = static_cast<const bf16*>(h), ga.ldh = ldh, ga.scale = scale, ga.colsum = dbias;
  ga.out = static_cast<bf16*>(dpre), ga.ldo = ldo;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool narrow = ceil_div(M, BM) * ceil_div(N, 256) < 100 || N % 256 != 0;
  return narrow ? launch_gemm<128, true, EPI_RELU_BWD>(dy, lddy, w, ldw, ga, st)
                : launch_gemm<256, true, EPI_RELU_BWD>(dy, lddy, w, ldw, ga, st);
}

extern "C" int destr_gemm_res_ln(const void* a, int lda, const void* w, int ldw, int M, int K, const float* bias,
                                 const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site, const void* res,
                                 int ldres, const float* gamma, const float* beta, void* z, int ldz, void* y, int ldy,
                                 float* mean, float* rstd, const void* res2, int ldres2, const float* gamma2,
                                 const float* beta2, void* y2, int ldy2, float* mean2, float* rstd2, void* stream) {
  int rc = check_common(a, lda, w, ldw, M, 256, K, 0);
  if (rc) return rc;
  DESTR_CHECK_ARG(res && gamma && beta && y && mean && rstd, "null pointer");
  DESTR_CHECK_ARG(ldres % 8 == 0 && ldy % 8 == 0 && ldres >= 256 && ldy >= 256 && (!z || (ldz % 8 == 0 && ldz >= 256)), "pitches");
  DESTR_CHECK_ARG(!res2 || (gamma2 && beta2 && y2 && mean2 && rstd2 && ldres2 % 8 == 0 && ldy2 % 8 == 0 && ldres2 >= 256 && ldy2 >= 256),
                  "second LayerNorm operands");
  GemmArgs ga{};
  ga.M = M, ga.N = 256, ga.K = K;
  ga.bias = bias, ga.dp = Drop{drop_seed, drop_thr16, drop_site};
  ga.res = static_cast<const bf16*>(res), ga.ldres = ldres, ga.gamma = gamma, ga.beta = beta;
  ga.z = static_cast<bf16*>(z), ga.ldz = ldz, ga.out = static_cast<bf16*>(y), ga.ldo = ldy;
  ga.mean = mean, ga.rstd = rstd;
  ga.res2 = static_cast<const bf16*>(res2), ga.ldres2 = ldres2, ga.gamma2 = gamma2, ga.beta2 = beta2;
  ga.y2 = static_cast<bf16*>(y2), ga.ldy2 = ldy2, ga.mean2 = mean2, ga.rstd2 = rstd2;
  return launch_gemm<256, false, EPI_RES_LN>(a, lda, w, ldw, ga, static_cast<cudaStream_t>(stream));
}

extern "C" int destr_gemm_dw(const void* dy, int lddy, const void* x, int ldx, int M, int Nout, int Kin, float* dw,
                             int lddw, void* stream) {
  DESTR_CHECK_ARG(dy && x && dw, "null pointer");
  DESTR_CHECK_ARG(M > 0 && Nout > 0 && Kin > 0 && Nout % 8 == 0 && Kin % 8 == 0, "shape (Nout % 8 == 0, Kin % 8 == 0)");
  DESTR_CHECK_ARG(lddy % 8 == 0 && ldx % 8 == 0 && lddy >= Nout && ldx >= Kin && lddw % 4 == 0 && lddw >= Kin, "pitches");
  DESTR_CHECK_ARG((reinterpret_cast<uintptr_t>(dw) & 15) == 0, "dw must be 16-byte aligned");
  CUtensorMap tdy, tx;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tdy, dy, M, Nout, lddy, BK, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tx, x, M, Kin, ldx, BK, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const size_t smem = sizeof(SmemDw) + 1024;
  DESTR_SMEM_OPTIN(gemm_dw_kernel, smem);
  const int tiles = ceil_div(Nout, BM) * ceil_div(Kin, DW_BN);
  const int nchunk = ceil_div(M, BK);
  // Split of the M rows: every split adds one fp32 red.global.add pass over the whole output (measured: 66 splits of
  // a 256 x 256 gradient = 17 MB of atomics for 0.26 MB of result, and the kernel was bound by them), so a CTA keeps
  // at least 8 chunks (512 rows) and the grid stays within two CTAs per SM.
  int S = (2 * 148) / tiles;
  if (S > nchunk / 8) S = nchunk / 8;
  if (g_dw_split > 0) S = g_dw_split;
  if (S < 1) S = 1;
  if (S > nchunk) S = nchunk;
  const int per = ceil_div(nchunk, S);
  S = ceil_div(nchunk, per);
  DESTR_CUDA(launch_k(gemm_dw_kernel, dim3(tiles, S), dim3(DW_THREADS), smem, static_cast<cudaStream_t>(stream), tdy, tx,
                      dw, lddw, Nout, Kin, nchunk, per));
  return 0;
}

extern "C" int destr_gemm_bf16_batched(const void* a, int lda, int a_batch_rows, const void* b, int ldb,
                                       int b_batch_rows, int b_kn, int batch, int M, int N, int K, void* out, int ldo,
                                       void* stream) {
  int rc = check_common(a, lda, b, ldb, M, N, K, b_kn);
  if (rc) return rc;
  DESTR_CHECK_ARG(out && ldo % 8 == 0 && ldo >= N && batch > 0 && a_batch_rows >= M && b_batch_rows > 0, "out / batch");
  GemmArgs ga{};
  ga.M = M, ga.N = N, ga.K = K;
  ga.a_brows = a_batch_rows, ga.b_brows = b_batch_rows;
  ga.out = static_cast<bf16*>(out), ga.ldo = ldo;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return b_kn ? launch_gemm<128, true, EPI_STORE>(a, lda, b, ldb, ga, st, batch)
              : launch_gemm<128, false, EPI_STORE>(a, lda, b, ldb, ga, st, batch);
}

// two batched products with the same (N, K, batch, b_kn) in ONE launch (gridDim.z = 2)
extern "C" int destr_gemm_bf16_batched2(const void* a1, int lda1, int a1_batch_rows, const void* b1, int ldb1,
                                        int b1_batch_rows, int M1, void* out1, int ldo1, const void* a2, int lda2,
                                        int a2_batch_rows, const void* b2, int ldb2, int b2_batch_rows, int M2,
                                        void* out2, int ldo2, int b_kn, int batch, int N, int K, void* stream) {
  int rc = check_common(a1, lda1, b1, ldb1, M1, N, K, b_kn);
  if (rc) return rc;
  if ((rc = check_common(a2, lda2, b2, ldb2, M2, N, K, b_kn))) return rc;
  DESTR_CHECK_ARG(out1 && out2 && ldo1 % 8 == 0 && ldo2 % 8 == 0 && ldo1 >= N && ldo2 >= N && batch > 0, "out / batch");
  DESTR_CHECK_ARG(a1_batch_rows >= M1 && a2_batch_rows >= M2 && b1_batch_rows > 0 && b2_batch_rows > 0, "batch rows");
  GemmArgs g1{}, g2{};
  g1.M = M1, g1.N = N, g1.K = K, g1.a_brows = a1_batch_rows, g1.b_brows = b1_batch_rows;
  g1.out = static_cast<bf16*>(out1), g1.ldo = ldo1;
  g2.M = M2, g2.N = N, g2.K = K, g2.a_brows = a2_batch_rows, g2.b_brows = b2_batch_rows;
  g2.out = static_cast<bf16*>(out2), g2.ldo = ldo2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return b_kn ? launch_gemm<128, true, EPI_STORE>(a1, lda1, b1, ldb1, g1, st, batch, a2, lda2, b2, ldb2, &g2)
              : launch_gemm<128, false, EPI_STORE>(a1, lda1, b1, ldb1, g1, st, batch, a2, lda2, b2, ldb2, &g2);
}
