// Linear sum assignment on the GPU, one warp per image, bit-identical to scipy.optimize.linear_sum_assignment
// (the call at reference src/utils/matcher.py:109-112,186-189: per-image Hungarian matching on the diagonal
// blocks of the cost matrix).  Removes the only host round trip of a training step (`C.cpu()` + scipy).
//
// The algorithm is scipy's (rectangular_lsap.cpp: Crouse's shortest augmenting path with dual variables),
// restated in oracle/lsap_oracle.py and followed here step by step IN DOUBLE PRECISION: the `remaining` list
// (filled in reverse, swap-removed), the scan order and the tie rule "strictly lower, or equal and unassigned"
// are reproduced exactly, so the assignment is identical even when costs tie.  The scan over the remaining
// columns is the parallel part: each lane keeps (lowest, first position with it, last unassigned position
// with it) over its strided share, and the warp combines them with the rule that is equivalent to the
// sequential scan:  lower value wins; on equal values first = min, last-unassigned = max.
#include <math_constants.h>

#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

struct Best {
  double lowest;
  int first, last_unassigned;
};
// combining rule of two partial scans (equivalent to the sequential scan): the lower value wins; on equal values
// first = min(first), last_unassigned = max(last_unassigned)
// warp-wide combine with four redux.sync instead of a 5-round shuffle butterfly of (double, int, int): the double is
// mapped to an order-preserving 64-bit key, whose high and low words are minimised one after the other; among the
// lanes that hold the minimum, first = min and last_unassigned = max
__device__ __forceinline__ Best warp_best(Best v) {
  const unsigned full = 0xffffffffu;
  const long long bits = __double_as_longlong(v.lowest + 0.0);  // (+0.0: -0 and +0 must map to one key)
  const unsigned long long key = static_cast<unsigned long long>(bits) ^
                                 (bits < 0 ? 0xffffffffffffffffull : 0x8000000000000000ull);
  const unsigned hi = static_cast<unsigned>(key >> 32), lo = static_cast<unsigned>(key);
  const unsigned mhi = __reduce_min_sync(full, hi);
  const unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xffffffffu);
  const bool mine = hi == mhi && lo == mlo;
  Best r;
  r.first = static_cast<int>(__reduce_min_sync(full, mine ? static_cast<unsigned>(v.first) : 0x7fffffffu));
  r.last_unassigned = __reduce_max_sync(full, mine ? v.last_unassigned : -1);
  const unsigned src = __ffs(__ballot_sync(full, mine)) - 1;
  r.lowest = __shfl_sync(full, v.lowest, src);
  return r;
}

// one warp per image; dynamic shared memory: 3 double[M] + 4 int[M] + 2 uint8[M], M = max(Q, max targets)
__global__ void __launch_bounds__(32)
lsap_kernel(const float* __restrict__ cost, const int32_t* __restrict__ offs, int Q, int n_slots, int M, int cost_in_smem,
            int64_t* __restrict__ pred_idx, int64_t* __restrict__ tgt_idx, uint8_t* __restrict__ valid,
            int32_t* __restrict__ status) {
  extern __shared__ double smem_d[];
  double* u = smem_d;
  double* v = u + M;
  double* spc = v + M;
  int* path = reinterpret_cast<int*>(spc + M);
  int* col4row = path + M;
  int* row4col = col4row + M;
  int* remaining = row4col + M;
  uint8_t* SR = reinterpret_cast<uint8_t*>(remaining + M);
  uint8_t* SC = SR + M;

  const int b = blockIdx.x, lane = threadIdx.x;
  const int T = offs[b + 1] - offs[b];
  const float* C = cost + static_cast<size_t>(Q) * offs[b];  // [Q, T] row-major
  int64_t* out_p = pred_idx + static_cast<size_t>(b) * n_slots;
  int64_t* out_t = tgt_idx + static_cast<size_t>(b) * n_slots;
  uint8_t* out_v = valid + static_cast<size_t>(b) * n_slots;
  for (int k = lane; k < n_slots; k += 32) {  // padded slots: query index Q, target 0, invalid
    out_p[k] = Q;
    out_t[k] = 0;
    out_v[k] = 0;
  }
  if (lane == 0) status[b] = 0;
  if (T <= 0) return;

  // scipy rejects NaN / -inf entries (ValueError): report instead of assigning
  int bad = 0;
  for (int e = lane; e < Q * T; e += 32) {
    const float c = C[e];
    bad |= (c != c) || (c == -CUDART_INF_F);
  }
  if (__any_sync(0xffffffffu, bad)) {
    if (lane == 0) status[b] = 1;
    return;
  }

  const bool transpose = T < Q;  // tall matrix: rows = targets, columns = queries
  const int nr = transpose ? T : Q, nc = transpose ? Q : T;
  // the scans re-read one row of the (possibly transposed) cost block per iteration: stage the block in shared
  // memory, row-major in the orientation of the solve (coalesced scans, no L2 latency on the serial path)
  float* Cs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(SC + M) + 15) & ~uintptr_t(15));
  const bool staged = cost_in_smem != 0;
  if (staged) {
    for (int e = lane; e < Q * T; e += 32) {
      const int q = e / T, t = e - q * T;
      Cs[transpose ? t * Q + q : e] = C[e];
    }
  }
  auto cst = [&](int r, int c) -> double {
    if (staged) return static_cast<double>(Cs[r * nc + c]);
    return static_cast<double>(transpose ? C[static_cast<size_t>(c) * T + r] : C[static_cast<size_t>(r) * T + c]);
  };
  for (int k = lane; k < nr; k += 32) {
    u[k] = 0.0;
    col4row[k] = -1;
  }
  for (int k = lane; k < nc; k += 32) {
    v[k] = 0.0;
    path[k] = -1;
    row4col[k] = -1;
  }
  __syncwarp();

  for (int cur = 0; cur < nr; ++cur) {
    for (int k = lane; k < nc; k += 32) {
      remaining[k] = nc - k - 1;
      SC[k] = 0;
      spc[k] = CUDART_INF;
    }
    for (int k = lane; k < nr; k += 32) SR[k] = 0;
    __syncwarp();
    int num_remaining = nc, sink = -1, i = cur;
    double min_val = 0.0;
    while (sink == -1) {
      if (lane == 0) SR[i] = 1;
      const double ui = u[i];
      Best best{CUDART_INF, 0x7fffffff, -1};
      // this lane's columns, four at a time: the (independent) loads and fp64 updates of the four are issued
      // together, then folded into `best` in scan order -- the serial chain per step is one column, not four
      for (int it0 = lane; it0 < num_remaining; it0 += 128) {
        int jj[4];
        double rr[4], sp[4];
        bool un[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int it = it0 + 32 * q;
          jj[q] = it < num_remaining ? remaining[it] : -1;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (jj[q] >= 0) {
            const int j = jj[q];
            rr[q] = min_val + cst(i, j) - ui - v[j];
            sp[q] = spc[j];
            un[q] = row4col[j] == -1;
          }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (jj[q] >= 0) {
            const int j = jj[q], it = it0 + 32 * q;
            double s = sp[q];
            if (rr[q] < s) {
              path[j] = i;
              spc[j] = rr[q];
              s = rr[q];
            }
            if (s < best.lowest) {
              best = Best{s, it, un[q] ? it : -1};
            } else if (s == best.lowest) {
              if (best.first == 0x7fffffff) best.first = it;  // (only when everything so far was +inf)
              if (un[q]) best.last_unassigned = it;
            }
          }
      }
      best = warp_best(best);
      min_val = best.lowest;
      if (min_val == CUDART_INF) {  // infeasible (cannot happen for finite costs)
        if (lane == 0) status[b] = 2;
        return;
      }
      const int index = best.last_unassigned >= 0 ? best.last_unassigned : best.first;
      int j = 0, nxt = 0;
      if (lane == 0) {
        j = remaining[index];
        nxt = row4col[j];
        SC[j] = 1;
        remaining[index] = remaining[num_remaining - 1];
      }
      j = __shfl_sync(0xffffffffu, j, 0);
      nxt = __shfl_sync(0xffffffffu, nxt, 0);
      --num_remaining;
      if (nxt == -1) sink = j; else i = nxt;
      __syncwarp();
    }
    // dual update
    if (lane == 0) u[cur] += min_val;
    for (int r = lane; r < nr; r += 32)
      if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
    for (int c = lane; c < nc; c += 32)
      if (SC[c]) v[c] -= min_val - spc[c];
    __syncwarp();
    // augment the previous solution along the path
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = path[j];
        row4col[j] = r;
        const int t = col4row[r];
        col4row[r] = j;
        j = t;
        if (r == cur) break;
      }
    }
    __syncwarp();
  }

  // output in ascending query order (scipy: rows ascending; for the transposed problem argsort(col4row))
  int base = 0;
  for (int q0 = 0; q0 < Q; q0 += 32) {
    const int q = q0 + lane;
    int t = -1;
    if (q < Q) t = transpose ? row4col[q] : col4row[q];
    const unsigned m = __ballot_sync(0xffffffffu, t >= 0);
    if (t >= 0) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      out_p[pos] = q;
      out_t[pos] = t;
      out_v[pos] = 1;
    }
    base += __popc(m);
  }
}

}  // namespace
}  // namespace destr

extern "C" int destr_lsap_blockdiag(const float* cost, const int32_t* tgt_offsets, int B, int Q, int max_targets,
                                    int n_slots, int64_t* pred_idx, int64_t* tgt_idx, uint8_t* valid,
                                    int32_t* status, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(cost && tgt_offsets && pred_idx && tgt_idx && valid && status, "null pointer");
  DESTR_CHECK_ARG(B > 0 && Q > 0 && max_targets > 0, "shape");
  DESTR_CHECK_ARG(n_slots >= (Q < max_targets ? Q : max_targets), "n_slots must be >= min(Q, max_targets)");
  const int M = Q > max_targets ? Q : max_targets;
  size_t smem = (static_cast<size_t>(M) * (3 * sizeof(double) + 4 * sizeof(int) + 2) + 15) / 16 * 16;
  DESTR_CHECK_ARG(smem <= 200 * 1024, "problem too large for one warp's shared memory");
  const size_t block = static_cast<size_t>(Q) * max_targets * sizeof(float);
  const int cost_in_smem = smem + block <= 200 * 1024 ? 1 : 0;
  if (cost_in_smem) smem += block;
  DESTR_SMEM_OPTIN(lsap_kernel, smem);
  lsap_kernel<<<B, 32, smem, static_cast<cudaStream_t>(stream)>>>(cost, tgt_offsets, Q, n_slots, M, cost_in_smem, pred_idx,
                                                                  tgt_idx, valid, status);
  DESTR_LAUNCH_CHECK();
  return 0;
}
