// Mask packing, sine positional embeddings and the bf16 elementwise kernels of the encoder
// (HBM-bound: 128-bit vector loads/stores, no shared memory, grid sized in multiples of 148 SMs).
#include "../../include/destr_b200.h"
#include "common.cuh"

namespace destr {
namespace {

constexpr int kSMs = 148;

__global__ void pack_key_mask_kernel(const uint8_t* __restrict__ kpm, uint32_t* __restrict__ bits, int n_keys,
                                     int words_per_row) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  const int b = blockIdx.y;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words_per_row) return;
  uint32_t word = 0;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    const int key = w * 32 + i;
    bool masked = key >= n_keys;
    if (!masked && kpm) masked = kpm[static_cast<size_t>(b) * n_keys + key] != 0;
    word |= (masked ? 1u : 0u) << i;
  }
  bits[static_cast<size_t>(b) * words_per_row + w] = word;
}

// 10000^(2*floor(i/2)/128), i = channel within a 128-wide half
__device__ __forceinline__ float sine_div(int i) { return powf(10000.0f, (2.0f * (float)(i >> 1)) / 128.0f); }

// one block (128 threads) per image ROW (b, y): the cumulative valid-pixel counts of the row and of every column
// up to it are computed once into shared memory (exact in fp32), then thread c writes channels c (y half) and
// 128+c (x half) of the row's W tokens.
constexpr int kMaxW = 512;
__global__ void __launch_bounds__(128)
sine_pos2d_kernel(const uint8_t* __restrict__ mask, float* __restrict__ pos_f32, __nv_bfloat16* __restrict__ pos_bf16,
                  int H, int W) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  __shared__ float s_ye[kMaxW], s_xe[kMaxW];
  const int b = blockIdx.x / H, y = blockIdx.x - b * H;
  const uint8_t* mb = mask + static_cast<size_t>(b) * H * W;
  const float two_pi = 6.283185307179586f;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    float ycum = 0.f, ytot = 0.f, xcum = 0.f, xtot = 0.f;
    for (int yy = 0; yy < H; ++yy) {
      const float v = mb[yy * W + x] ? 0.f : 1.f;
      ytot += v;
      if (yy <= y) ycum += v;
    }
    for (int xx = 0; xx < W; ++xx) {
      const float v = mb[y * W + xx] ? 0.f : 1.f;
      xtot += v;
      if (xx <= x) xcum += v;
    }
    s_ye[x] = __fmul_rn(__fdiv_rn(ycum, __fadd_rn(ytot, 1e-6f)), two_pi);
    s_xe[x] = __fmul_rn(__fdiv_rn(xcum, __fadd_rn(xtot, 1e-6f)), two_pi);
  }
  __syncthreads();
  const int c = threadIdx.x;
  const float dv = sine_div(c);
  for (int x = 0; x < W; ++x) {
    const float ay = __fdiv_rn(s_ye[x], dv), ax = __fdiv_rn(s_xe[x], dv);
    const float vy = (c & 1) ? cosf(ay) : sinf(ay);
    const float vx = (c & 1) ? cosf(ax) : sinf(ax);
    const size_t o = (static_cast<size_t>(blockIdx.x) * W + x) * 256;
    if (pos_f32) {
      pos_f32[o + c] = vy;
      pos_f32[o + 128 + c] = vx;
    }
    if (pos_bf16) {
      pos_bf16[o + c] = __float2bfloat16(vy);
      pos_bf16[o + 128 + c] = __float2bfloat16(vx);
    }
  }
}

__global__ void query_sine_embed_kernel(const float* __restrict__ centers, float* __restrict__ out_f32,
                                        __nv_bfloat16* __restrict__ out_bf16, int M) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  const int r = blockIdx.x;
  const int c = threadIdx.x;  // 0..127
  const float two_pi = 6.283185307179586f;
  const float xe = __fmul_rn(centers[2 * r + 0], two_pi);
  const float ye = __fmul_rn(centers[2 * r + 1], two_pi);
  const float dv = sine_div(c);
  const float ay = __fdiv_rn(ye, dv), ax = __fdiv_rn(xe, dv);
  const float vy = (c & 1) ? cosf(ay) : sinf(ay);
  const float vx = (c & 1) ? cosf(ax) : sinf(ax);
  const size_t o = static_cast<size_t>(r) * 256;
  if (out_f32) {
    out_f32[o + c] = vy;
    out_f32[o + 128 + c] = vx;
  }
  if (out_bf16) {
    out_bf16[o + c] = __float2bfloat16(vy);
    out_bf16[o + 128 + c] = __float2bfloat16(vx);
  }
}

// ---- bf16 x8 vector helpers ----
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  F8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f.v[2 * i], f.v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// MODE 0: y = x + pos*s ; 1: ds = dy*pos (x unused) ; 2: y = a*b (s unused)
template <int MODE>
__global__ void ew3_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ p,
                           const __nv_bfloat16* __restrict__ s, __nv_bfloat16* __restrict__ y, int64_t n8) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    F8 r;
    if (MODE == 0) {
      const F8 a = ld8(x + i * 8), b = ld8(p + i * 8), c = ld8(s + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) r.v[k] = fmaf(b.v[k], c.v[k], a.v[k]);
    } else {
      const F8 a = ld8(x + i * 8), b = ld8(p + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) r.v[k] = a.v[k] * b.v[k];
    }
    st8(y + i * 8, r);
  }
}

// MODE 3: dx_out = dx_in + dy ; ds = dy*pos   (backward of y = x + pos*s with the residual gradient folded in)
__global__ void pos_mul_add_bwd_acc_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ pos,
                                           const __nv_bfloat16* __restrict__ dx_in, __nv_bfloat16* __restrict__ ds,
                                           __nv_bfloat16* __restrict__ dx_out, int64_t n8) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const F8 g = ld8(dy + i * 8), p = ld8(pos + i * 8), r = ld8(dx_in + i * 8);
    F8 a, b;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a.v[k] = g.v[k] * p.v[k];
      b.v[k] = g.v[k] + r.v[k];
    }
    st8(ds + i * 8, a);
    st8(dx_out + i * 8, b);
  }
}

// dpre = dy * (h > 0); dbias[c] += sum_rows dpre[:, c].  Block = 32 column groups (8 channels each) x 8 row
// lanes; a block owns 32*CHUNKS rows of a 256-column stripe, walks them in chunks of 32 rows and issues the 4 row
// loads of a chunk up front (latency-bound otherwise); ONE atomicAdd per column per block at the end, so large
// matrices use CHUNKS = 4 (4x fewer same-address atomics, which bounded the FFN-sized launches).
// grid = (ceil(C/256), ceil(M/(32*CHUNKS))).
constexpr int kRowsPerChunk = 32;
template <int CHUNKS>
__global__ void __launch_bounds__(256)
relu_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ h,
                       __nv_bfloat16* __restrict__ dpre, float* __restrict__ dbias, int M, int C, int lddy, int ldh,
                       int ldo, float scale) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  __shared__ float red[8][256];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (col < C) {
#pragma unroll 1
    for (int ch = 0; ch < CHUNKS; ++ch) {
      const int row0 = (blockIdx.y * CHUNKS + ch) * kRowsPerChunk + rl;
      F8 g[kRowsPerChunk / 8], a[kRowsPerChunk / 8];
#pragma unroll
      for (int u = 0; u < kRowsPerChunk / 8; ++u) {
        const int row = row0 + u * 8;
        if (row < M) {
          g[u] = ld8(dy + (size_t)row * lddy + col);
          if (h) a[u] = ld8(h + (size_t)row * ldh + col);
        }
      }
#pragma unroll
      for (int u = 0; u < kRowsPerChunk / 8; ++u) {
        const int row = row0 + u * 8;
        if (row < M) {
          if (h) {
#pragma unroll
            for (int k = 0; k < 8; ++k) g[u].v[k] = a[u].v[k] > 0.f ? g[u].v[k] * scale : 0.f;
            st8(dpre + (size_t)row * ldo + col, g[u]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += g[u].v[k];
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cg * 8 + k] = acc[k];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < C) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][c];
    atomicAdd(dbias + blockIdx.x * 256 + c, s);
  }
}

// x <- dropout(x) in place (the dropout after a fused GEMM+ReLU: encoder_block.py:108 dropout2, decoder_block.py:255):
// thread = 8 consecutive channels of one row, one hash per channel pair
__global__ void __launch_bounds__(256)
dropout_inplace_kernel(__nv_bfloat16* __restrict__ x, int M, int C, int ld, Drop dp) {
  pdl_wait();  // (common.cuh: programmatic dependent launch) nothing global is touched before this
  pdl_launch();
  const uint32_t seed = dp.seed ? *dp.seed : 0u;
  const float s = drop_scale(dp.thr16);
  const int c8 = C / 8;
  const int64_t n = static_cast<int64_t>(M) * c8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / c8), col = static_cast<int>(i - static_cast<int64_t>(row) * c8) * 8;
    F8 v = ld8(x + static_cast<size_t>(row) * ld + col);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t bits = drop_bits(seed, dp.site, row, (col >> 1) + k);
      v.v[2 * k] = ((bits & 0xFFFFu) >= dp.thr16) ? v.v[2 * k] * s : 0.f;
      v.v[2 * k + 1] = ((bits >> 16) >= dp.thr16) ? v.v[2 * k + 1] * s : 0.f;
    }
    st8(x + static_cast<size_t>(row) * ld + col, v);
  }
}

inline int ew_grid(int64_t n8, int threads) {
  int64_t blocks = (n8 + threads - 1) / threads;
  const int64_t cap = kSMs * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace destr

using namespace destr;

extern "C" int destr_pack_key_mask(const uint8_t* kpm, uint32_t* bits, int B, int n_keys, int words_per_row,
                                   void* stream) {
  DESTR_CHECK_ARG(bits && B > 0 && n_keys > 0 && words_per_row * 32 >= n_keys, "shape");
  dim3 grid(ceil_div(words_per_row, 64), B);
  DESTR_CUDA(launch_k(pack_key_mask_kernel, dim3(grid), dim3(64), 0, (cudaStream_t)stream, kpm, bits, n_keys, words_per_row));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_sine_pos2d(const uint8_t* mask, float* pos_f32, void* pos_bf16, int B, int H, int W,
                                void* stream) {
  DESTR_CHECK_ARG(mask && (pos_f32 || pos_bf16) && B > 0 && H > 0 && W > 0 && W <= kMaxW, "shape (W <= 512)");
  DESTR_CUDA(launch_k(sine_pos2d_kernel, dim3(B * H), dim3(128), 0, (cudaStream_t)stream, mask, pos_f32, (__nv_bfloat16*)pos_bf16, H, W));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_query_sine_embed(const float* centers, float* out_f32, void* out_bf16, int M, void* stream) {
  DESTR_CHECK_ARG(centers && (out_f32 || out_bf16) && M > 0, "shape");
  DESTR_CUDA(launch_k(query_sine_embed_kernel, dim3(M), dim3(128), 0, (cudaStream_t)stream, centers, out_f32, (__nv_bfloat16*)out_bf16, M));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_pos_mul_add_fwd(const void* x, const void* pos, const void* s, void* y, int64_t n_elem,
                                     void* stream) {
  DESTR_CHECK_ARG(x && pos && s && y && n_elem > 0 && n_elem % 8 == 0, "n_elem must be a multiple of 8");
  const int64_t n8 = n_elem / 8;
  DESTR_CUDA(launch_k(ew3_kernel<0>, dim3(ew_grid(n8, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)pos, (const __nv_bfloat16*)s, (__nv_bfloat16*)y, n8));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_pos_mul_add_bwd(const void* dy, const void* pos, void* ds, int64_t n_elem, void* stream) {
  DESTR_CHECK_ARG(dy && pos && ds && n_elem > 0 && n_elem % 8 == 0, "n_elem must be a multiple of 8");
  const int64_t n8 = n_elem / 8;
  DESTR_CUDA(launch_k(ew3_kernel<1>, dim3(ew_grid(n8, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)pos, nullptr, (__nv_bfloat16*)ds, n8));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_mul_fwd(const void* a, const void* b, void* y, int64_t n_elem, void* stream) {
  DESTR_CHECK_ARG(a && b && y && n_elem > 0 && n_elem % 8 == 0, "n_elem must be a multiple of 8");
  const int64_t n8 = n_elem / 8;
  DESTR_CUDA(launch_k(ew3_kernel<2>, dim3(ew_grid(n8, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, nullptr, (__nv_bfloat16*)y, n8));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_pos_mul_add_bwd_acc(const void* dy, const void* pos, const void* dx_in, void* ds, void* dx_out,
                                         int64_t n_elem, void* stream) {
  DESTR_CHECK_ARG(dy && pos && dx_in && ds && dx_out && n_elem > 0 && n_elem % 8 == 0, "n_elem must be a multiple of 8");
  const int64_t n8 = n_elem / 8;
  DESTR_CUDA(launch_k(pos_mul_add_bwd_acc_kernel, dim3(ew_grid(n8, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)pos, (const __nv_bfloat16*)dx_in, (__nv_bfloat16*)ds,
      (__nv_bfloat16*)dx_out, n8));
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_relu_bwd_colsum(const void* dy, int lddy, const void* h, int ldh, void* dpre, int ldo,
                                     float* dbias, int M, int C, float scale, void* stream) {
  DESTR_CHECK_ARG(dy && dbias && M > 0 && C > 0 && C % 8 == 0 && lddy % 8 == 0, "shape");
  DESTR_CHECK_ARG((h == nullptr) == (dpre == nullptr), "h and dpre go together (both NULL = plain column sum)");
  if (static_cast<int64_t>(M) * C >= (1 << 22)) {  // FFN-sized: fewer, longer blocks (atomics on dbias bound the short ones)
    dim3 grid(ceil_div(C, 256), ceil_div(M, kRowsPerChunk * 4));
    DESTR_CUDA(launch_k(relu_bwd_colsum_kernel<4>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)h, (__nv_bfloat16*)dpre, dbias, M, C, lddy, ldh, ldo, scale));
  } else {
    dim3 grid(ceil_div(C, 256), ceil_div(M, kRowsPerChunk));
    DESTR_CUDA(launch_k(relu_bwd_colsum_kernel<1>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)h, (__nv_bfloat16*)dpre, dbias, M, C, lddy, ldh, ldo, scale));
  }
  DESTR_LAUNCH_CHECK();
  return 0;
}

extern "C" int destr_dropout_inplace(void* x, int ld, int M, int C, const uint32_t* drop_seed, uint32_t drop_thr16,
                                     uint32_t drop_site, void* stream) {
  DESTR_CHECK_ARG(x && M > 0 && C > 0 && C % 8 == 0 && ld % 8 == 0 && ld >= C, "shape");
  if (drop_thr16 == 0) return 0;
  const int64_t n = static_cast<int64_t>(M) * (C / 8);
  DESTR_CUDA(launch_k(dropout_inplace_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (__nv_bfloat16*)x, M, C, ld, destr::Drop{drop_seed, drop_thr16, drop_site}));
  DESTR_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// Batch hand-over: up to 16 device-to-device copies in ONE launch.  The engine moves a staged batch (features, mask,
// selected queries, centres, packed targets: 9 small buffers) into the static inputs of the step graph before every
// replay; nine back-to-back memcpy nodes cost ~5 us each on the stream the graph replays on.
// ------------------------------------------------------------------------------------------------------
namespace destr {
namespace {
struct CopyMany {
  const uint8_t* src[16];
  uint8_t* dst[16];
  long long bytes[16];
  int n;
};
__global__ void copy_many_kernel(const CopyMany cm) {
  for (int k = blockIdx.y; k < cm.n; k += gridDim.y) {
    const uint8_t* __restrict__ s = cm.src[k];
    uint8_t* __restrict__ d = cm.dst[k];
    const long long nb = cm.bytes[k];
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
      const long long n16 = nb >> 4;
      for (long long i = tid; i < n16; i += nth) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
      for (long long i = (n16 << 4) + tid; i < nb; i += nth) d[i] = s[i];
    } else {
      for (long long i = tid; i < nb; i += nth) d[i] = s[i];
    }
  }
}
}  // namespace
}  // namespace destr

extern "C" int destr_copy_many(const void* const* src, void* const* dst, const int64_t* bytes, int n, void* stream) {
  using namespace destr;
  DESTR_CHECK_ARG(src && dst && bytes && n > 0 && n <= 16, "1..16 buffers (HOST arrays of device pointers / sizes)");
  CopyMany cm{};
  long long big = 0;
  for (int k = 0; k < n; ++k) {
    DESTR_CHECK_ARG(src[k] && dst[k] && bytes[k] >= 0, "null buffer / negative size");
    cm.src[k] = static_cast<const uint8_t*>(src[k]);
    cm.dst[k] = static_cast<uint8_t*>(dst[k]);
    cm.bytes[k] = bytes[k];
    big = bytes[k] > big ? bytes[k] : big;
  }
  cm.n = n;
  int bx = static_cast<int>((big / 16 + 255) / 256);
  bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
  copy_many_kernel<<<dim3(bx, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(cm);
  DESTR_LAUNCH_CHECK();
  return 0;
}

