// Bit matrices of the encoder attention-probability dropout mask (nn.MultiheadAttention(dropout=p),
// reference encoder_block.py:58-60), in the two orientations the flash kernels consume.
//
// The mask itself is the counter-based function of common.cuh (drop_bits / drop_keep; oracle/dropout_mask.py is the
// numpy twin): element (row = (b,h,query), col = key) is dropped when its 16-bit field < thr16.  Evaluating that hash
// inside the attention kernels costs 7 (forward) to 13 (backward, transposed access) integer instructions per score in
// loops that are already issue-bound; here ONE hash per key pair is evaluated once per layer by a streaming kernel and
// the decisions are stored as bits (1 = dropped), word-major so that the 32 lanes of a warp (consecutive query rows in
// the forward kernel, consecutive key rows in the backward one) read consecutive words:
//   rowbits[bh][w][q]   bit i <-> key   32 w + i    (forward: a thread owns a query row, consumes 48 keys per tile)
//   colbits[bh][w][k]   bit i <-> query 32 w + i    (backward: a thread owns a key row, consumes 32 queries per sub-tile)
// Np = 128 ceil(N / 128) entries per word row, `words` word rows per (b,h) (>= Np / 32 and >= 3 ceil(N / 96)); only
// the ceil(N/32) x ceil(N/32) blocks that hold a real (query, key) are written -- everything else is read by the
// kernels only for padded queries / masked keys, whose probabilities are zero whatever the bit.
// A warp owns a 32 x 32 block: lane = query, 16 hashes give its row word (sign bits funnel-shifted in), and a 5-step
// shuffle butterfly transposes the block into the column words.
#include "common.cuh"
#include "../../include/destr_b200.h"

namespace destr {
namespace {

constexpr int kWordsPerWarp = 3;

__global__ void __launch_bounds__(128) attn_dropout_bits_kernel(const uint32_t* __restrict__ seed_ptr, uint32_t thr16,
                                                                uint32_t site0, uint32_t site_stride, int BH, int N,
                                                                int Np, int words, uint32_t* __restrict__ rowbits,
                                                                uint32_t* __restrict__ colbits) {
  const int lane = threadIdx.x & 31;
  const int nW = (N + 31) >> 5;
  const int qb = blockIdx.x * 4 + (threadIdx.x >> 5);  // block of 32 queries
  if (qb >= nW) return;
  const int sidx = blockIdx.z / BH, bh = blockIdx.z - sidx * BH;
  const uint32_t site = site0 + sidx * site_stride;
  const uint32_t seed = seed_ptr ? *seed_ptr : 0u;
  const int q = qb * 32 + lane;
  // the part of the hash that does not depend on the key pair
  const uint32_t h0 = drop_base(seed, site) ^ ((static_cast<uint32_t>(bh) * N + q) * 0x85EBCA77u);
  const size_t plane = static_cast<size_t>(blockIdx.z) * words * Np;
  const int kb_end = min(nW, (static_cast<int>(blockIdx.y) + 1) * kWordsPerWarp);
  for (int kb = blockIdx.y * kWordsPerWarp; kb < kb_end; ++kb) {
    uint32_t x = 0;
#pragma unroll
    for (int kp = 15; kp >= 0; --kp) {  // highest key first: each decision is shifted in at bit 0
      uint32_t h = h0 ^ (static_cast<uint32_t>(kb * 16 + kp) * 0xC2B2AE3Du);
      h ^= h >> 16;
      h *= 0x7FEB352Du;
      h ^= h >> 15;
      h *= 0x846CA68Bu;
      h ^= h >> 16;
      // field < thr16  <=>  sign bit of (field - thr16)  (both < 2^16)
      x = __funnelshift_l((h >> 16) - thr16, x, 1);
      x = __funnelshift_l((h & 0xFFFFu) - thr16, x, 1);
    }
    if (rowbits) rowbits[plane + static_cast<size_t>(kb) * Np + q] = x;
    if (colbits) {
      // 32 x 32 bit transpose across the warp: lane r holds row r -> lane c holds column c
#pragma unroll
      for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u
                                                                                                          : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
      }
      colbits[plane + static_cast<size_t>(qb) * Np + kb * 32 + lane] = x;
    }
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_attn_dropout_bits(const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site0,
                                       uint32_t drop_site_stride, int n_sites, int BH, int N, int words,
                                       uint32_t* rowbits, uint32_t* colbits, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(BH > 0 && N > 0 && n_sites > 0 && (rowbits || colbits), "shape / null pointer");
  const int Np = (N + 127) / 128 * 128, nW = (N + 31) / 32;
  DESTR_CHECK_ARG(words >= Np / 32 && words >= (N + 95) / 96 * 3, "words (need >= 4 per 128 queries and 3 per 96 keys)");
  DESTR_CHECK_ARG(static_cast<int64_t>(n_sites) * BH <= 65535, "n_sites * BH");
  const dim3 grid((nW + 3) / 4, (nW + kWordsPerWarp - 1) / kWordsPerWarp, n_sites * BH);
  attn_dropout_bits_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      drop_seed, drop_thr16, drop_site0, drop_site_stride, BH, N, Np, words, rowbits, colbits);
  DESTR_LAUNCH_CHECK();
  return 0;
}
